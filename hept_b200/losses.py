"""The tracking task's loss on the library's kernels: ``InfoNCELoss`` of the reference
(src/utils/losses.py:8-63) with the same constructor and ``forward(x, point_pairs, cluster_ids, recons, pts)`` surface.

Everything runs in libhept_sm100.so (``hept_infonce_fwd`` / ``hept_infonce_bwd``, csrc/loss.cu), forward and backward,
deterministically; there is no CPU path.  One deliberate difference from the reference: its per-point sum of negative-pair
terms comes back COMPACTED (one entry per point that owns a negative pair) and is then indexed with raw point numbers
(losses.py:48-51) — identical whenever every point up to the largest first index owns a negative pair (true for the
radius-graph pairs the dataset ships), undefined otherwise; here the sums are indexed by point number.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops


class _InfoNCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, point_pairs, cluster_ids, recons, pts, metric: str, tau: float):
        x = x.contiguous()
        pairs = point_pairs.contiguous()
        loss, saved = ops.infonce_fwd(x, pairs, cluster_ids.contiguous(), recons.float().contiguous(), pts.float().contiguous(),
                                      metric, tau)
        ctx.save_for_backward(x, pairs, saved)
        ctx.cfg = (metric, tau)
        return loss

    @staticmethod
    def backward(ctx, grad_loss):
        x, pairs, saved = ctx.saved_tensors
        metric, tau = ctx.cfg
        return ops.infonce_bwd(x, pairs, saved, grad_loss.float().contiguous(), metric, tau), None, None, None, None, None, None


class InfoNCELoss(nn.Module):
    def __init__(self, tau, dist_metric):
        super().__init__()
        if dist_metric not in ops.METRICS:
            raise NotImplementedError(dist_metric)
        self.tau = tau
        self.dist_metric = dist_metric

    def forward(self, x, point_pairs, cluster_ids, recons, pts, **kwargs):
        if not x.is_cuda:
            raise RuntimeError("hept_b200.InfoNCELoss runs on CUDA (sm_100a) only; there is no CPU path")
        return _InfoNCE.apply(x.float(), point_pairs.long(), cluster_ids.long(), recons, pts, self.dist_metric, float(self.tau))

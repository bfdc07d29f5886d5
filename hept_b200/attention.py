"""Drop-in ``HEPTAttention``: the reference's module surface over the sm_100a library.

Mirrors ``HEPTAttention`` of the reference (Graph-COM/HEPT):
  * ``example/hept.py:31-81``  — batched flavour, kwargs ``w_rpe, coords, combined_shifts``;
  * ``src/models/attention/hept.py:59-117`` — single-event flavour, kwargs
    ``w_rpe, coords, raw_size, regions_h, region_indices``.
Same constructor (``HEPTAttention(hash_dim, **kwargs)`` reading ``h_dim, num_heads, block_size,
n_hashes, num_w_per_dist`` and ignoring the rest), same ``forward(query, key, value, **kwargs)``
returning ``(N, h_dim)``, same state_dict keys (``out_linear.weight``, ``out_linear.bias``,
``e2lsh.alpha``, plus ``e2lsh.beta`` when a src/ checkpoint carries it), so
``load_state_dict(strict=True)`` works on both checkpoint flavours.

Everything from the q/k/v inputs to the module output runs in libhept_sm100.so: one native call for
a3..a12 forward and one for its backward, and ``out_linear`` (whose parameters stay in an ``nn.Linear``
for state_dict compatibility) on the library's streaming kernels, forward and backward.  There is no
CPU path: CPU tensors raise, a missing library raises.

Documented deviation: the src/ reference zeroes ``value[raw_size:]`` IN the caller's tensor
(src/models/attention/hept.py:91 acts on a view).  Here padding rows are treated as zero inside the
kernels and the caller's tensor is left untouched.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import ops


class E2LSH(nn.Module):
    """Holder of the frozen projection matrix (example/hept_utils.py:38-47; hash_utils.py:339-350)."""

    def __init__(self, n_hashes: int, n_heads: int, dim: int, r: float = 1.0, with_beta: bool = False):
        super().__init__()
        self.alpha = nn.Parameter(torch.normal(0, 1, (n_heads, dim, n_hashes)), requires_grad=False)
        if with_beta:  # unused by the forward pass; present in src/ checkpoints (hash_utils.py:344)
            self.beta = nn.Parameter(torch.rand(1, n_hashes) * r, requires_grad=False)

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        key = prefix + "beta"
        if key in state_dict and not hasattr(self, "beta"):
            self.beta = nn.Parameter(torch.empty_like(state_dict[key]), requires_grad=False)
        elif key not in state_dict and hasattr(self, "beta"):
            del self.beta
        return super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)


class _HeptCore(torch.autograd.Function):
    """query, key, value, w_rpe.weight -> out_pre (N, H*D); everything the reference does before out_linear."""

    @staticmethod
    def forward(ctx, query, key, value, w_rpe_weight, alpha, coords, combined_shifts, region_eta, region_phi,
                regions_h, dims: ops.Dims, K: int):
        q, k, v = query.contiguous(), key.contiguous(), value.contiguous()
        coords = coords.contiguous()
        regions = None if region_eta is None else (region_eta, region_phi)
        out_pre, den, scale, pos = ops.attention_fwd(dims, q, k, v, coords, w_rpe_weight.contiguous(), K, alpha,
                                                     combined_shifts=combined_shifts, region_indices=regions,
                                                     regions_h=regions_h)
        ctx.save_for_backward(q, k, v, coords, w_rpe_weight, scale, pos, out_pre, den)
        ctx.dims, ctx.K = dims, K
        return out_pre

    @staticmethod
    def backward(ctx, d_out_pre):
        q, k, v, coords, w, scale, pos, out_pre, den = ctx.saved_tensors
        d = ctx.dims
        dq, dk, dv, dscale = ops.attention_bwd(d, q, k, v, coords, scale, pos, out_pre, den, d_out_pre.contiguous())
        dw = None
        if ctx.needs_input_grad[3]:
            dw = ops.coord_scale_backward(w, scale, dscale, d.H, d.D, ctx.K)
        return dq, dk, dv, dw, None, None, None, None, None, None, None, None


class _AttnFront(torch.autograd.Function):
    """x (N, D) -> q, k, v (N, H*D): ``norm1`` then ``w_q / w_k / w_v`` of the reference's Attn block
    (example/transformer.py:157-158; src/models/baselines/transformer.py:209-212) on the library's kernels
    (csrc/attn_block.cu).  The gradient w.r.t. x is the one through norm1; the block's residual path stays with autograd."""

    @staticmethod
    def forward(ctx, x, norm_weight, norm_bias, w_q, w_k, w_v, H: int, D: int, eps: float):
        x = x.contiguous()
        q, k, v, xn, wt = ops.attn_qkv_fwd(x, norm_weight, norm_bias, w_q, w_k, w_v, H, D, eps)
        ctx.save_for_backward(x, xn, norm_weight, wt)
        ctx.shape = (H, D, eps)
        return q, k, v

    @staticmethod
    def backward(ctx, dq, dk, dv):
        x, xn, gamma, wt = ctx.saved_tensors
        H, D, eps = ctx.shape
        dx, dgam, dbet, dwq, dwk, dwv = ops.attn_qkv_bwd(x, xn, gamma, wt, dq.contiguous(), dk.contiguous(), dv.contiguous(), H, D, eps)
        return dx, dgam, dbet, dwq, dwk, dwv, None, None, None


def attn_front(x: torch.Tensor, norm1: nn.LayerNorm, w_q: nn.Linear, w_k: nn.Linear, w_v: nn.Linear, num_heads: int):
    """``norm1(x)`` followed by the three bias-free projections -> q, k, v, in one native forward (and one native backward)."""
    d = x.shape[-1]
    if not ops.attn_qkv_supported(num_heads, d):
        raise NotImplementedError(f"attn_front: (H={num_heads}, D={d}) not compiled in (hept_attn_qkv_supported)")
    return _AttnFront.apply(x.float(), norm1.weight, norm1.bias, w_q.weight, w_k.weight, w_v.weight, num_heads, d, norm1.eps)


class _OutLinear(torch.autograd.Function):
    """out_pre (N, H*D), weight (D, H*D), bias (D) -> out (N, D): ``out_linear`` of example/hept.py:80 on the library's own
    streaming kernels (csrc/out_linear.cu) instead of three library GEMMs."""

    @staticmethod
    def forward(ctx, out_pre, weight, bias, dims: ops.Dims):
        w, b = weight.contiguous(), bias.contiguous()
        ctx.save_for_backward(out_pre, w)
        ctx.dims = dims
        return ops.out_linear_fwd(dims, out_pre, w, b)

    @staticmethod
    def backward(ctx, d_out):
        out_pre, w = ctx.saved_tensors
        dx, dw, db = ops.out_linear_bwd(ctx.dims, d_out.contiguous(), w, out_pre, need_input_grad=ctx.needs_input_grad[0])
        return dx, dw, db, None


class HEPTAttention(nn.Module):
    def __init__(self, hash_dim: int, **kwargs):
        super().__init__()
        self.dim_per_head = kwargs["h_dim"]
        self.num_heads = kwargs["num_heads"]
        self.out_linear = nn.Linear(self.num_heads * self.dim_per_head, self.dim_per_head)
        self.block_size = kwargs["block_size"]
        self.n_hashes = kwargs["n_hashes"]
        self.num_w_per_dist = kwargs["num_w_per_dist"]
        self.hash_dim = hash_dim
        self.e2lsh = E2LSH(n_hashes=self.n_hashes, n_heads=self.num_heads, dim=hash_dim,
                           with_beta=bool(kwargs.get("e2lsh_beta", False)))

    # -- the hot path ---------------------------------------------------------------------------
    def attend(self, query, key, value, **kwargs) -> torch.Tensor:
        """Everything up to (not including) out_linear -> (N, H*D)."""
        return self._attend(query, key, value, **kwargs)[0]

    def _attend(self, query, key, value, **kwargs):
        if not query.is_cuda:
            raise RuntimeError(
                "hept_b200.HEPTAttention runs on CUDA (sm_100a) only; there is no CPU path. "
                "Use the reference module for CPU-side tooling (FLOP counters, tracing)."
            )
        n = query.shape[0]
        H, D = self.num_heads, self.dim_per_head
        coords = kwargs["coords"]
        C = coords.shape[-1]
        if D + C != self.hash_dim:
            raise ValueError(f"coords has {C} columns but hash_dim={self.hash_dim} implies {self.hash_dim - D}")
        if n % self.block_size != 0:
            raise ValueError(f"N={n} is not a multiple of block_size={self.block_size} (pad with prepare_input first)")
        w = kwargs["w_rpe"].weight
        # the same codes as int32 when the caller's prepare_input provides them (hept_b200.prepare does): 96 bytes per hit less
        shifts = kwargs.get("combined_shifts32")
        if shifts is None:
            shifts = kwargs.get("combined_shifts")
        eta = phi = regions_h = None
        raw = n
        if shifts is None:
            if "raw_size" not in kwargs:
                raise KeyError("need either `combined_shifts` (example/ flavour) or `raw_size`, `regions_h`, "
                               "`region_indices` (src/ flavour)")
            raw = int(kwargs["raw_size"])
            eta, phi = kwargs["region_indices"]
            regions_h = kwargs["regions_h"].contiguous()
            eta, phi = eta.contiguous(), phi.contiguous()
        dims = ops.Dims(N=n, H=H, D=D, C=C, T=self.n_hashes, B=self.block_size, raw_size=raw)
        f32 = lambda t: t if t.dtype == torch.float32 else t.float()
        out_pre = _HeptCore.apply(f32(query).reshape(n, H * D), f32(key).reshape(n, H * D), f32(value).reshape(n, H * D),
                                  w, self.e2lsh.alpha, f32(coords), shifts, eta, phi, regions_h, dims,
                                  self.num_w_per_dist)
        return out_pre, dims

    def forward(self, query, key, value, **kwargs):
        out_pre, dims = self._attend(query, key, value, **kwargs)
        return _OutLinear.apply(out_pre, self.out_linear.weight, self.out_linear.bias, dims)

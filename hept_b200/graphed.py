"""CUDA-graph replay of the hot path for fixed shapes (small events are launch-bound: a tracking-6k attention call is
~20 kernel launches of ~12 us each behind ~0.5 ms of host work per forward + backward).

The library never allocates or synchronises (include/hept_b200.h) and every output / workspace comes from torch's
allocator, so a whole module call — and its backward — can be captured into CUDA graphs with
``torch.cuda.make_graphed_callables`` and replayed with two graph launches per step.  Same kernels, same order, same
bits as the eager call (tests/test_graphed.py).  Shapes, the flavour of the kwargs and which tensors require gradients
are frozen at capture time; call the eager module for anything else.
"""
from __future__ import annotations

from typing import Sequence

import torch
import torch.nn as nn

from .attention import HEPTAttention


class _ExampleStep(nn.Module):
    """(query, key, value, coords, combined_shifts) -> module output; parameters = the module's and w_rpe's."""

    def __init__(self, attn: HEPTAttention, w_rpe: nn.Module):
        super().__init__()
        self.attn, self.w_rpe = attn, w_rpe

    def forward(self, query, key, value, coords, combined_shifts):
        return self.attn(query, key, value, w_rpe=self.w_rpe, coords=coords, combined_shifts=combined_shifts)


def graphed_attention(attn: HEPTAttention, w_rpe: nn.Module, query, key, value, coords, combined_shifts,
                      num_warmup_iters: int = 3):
    """-> callable(query, key, value, coords, combined_shifts) replaying captured forward / backward graphs of
    ``attn(query, key, value, w_rpe=w_rpe, coords=coords, combined_shifts=combined_shifts)`` (example/ flavour).
    The sample tensors fix shapes, dtypes and ``requires_grad`` flags."""
    if not query.is_cuda:
        raise RuntimeError("graphed_attention needs CUDA tensors (there is no CPU path)")
    step = _ExampleStep(attn, w_rpe)
    sample = tuple(t.detach().clone().requires_grad_(t.requires_grad) for t in (query, key, value)) + \
        (coords.detach().clone(), combined_shifts.detach().clone())
    with torch.cuda.device(query.device):
        return torch.cuda.make_graphed_callables(step, sample, num_warmup_iters=num_warmup_iters,
                                                 allow_unused_input=True)   # w_rpe.bias, e2lsh.alpha never get a gradient


class _BlockStep(nn.Module):
    """(x, coords, combined_shifts) -> attention output of an Attn block: norm1 -> w_q / w_k / w_v -> HEPTAttention."""

    def __init__(self, attn: HEPTAttention, w_rpe: nn.Module, norm1: nn.LayerNorm, w_q: nn.Linear, w_k: nn.Linear, w_v: nn.Linear):
        super().__init__()
        self.attn, self.w_rpe, self.norm1, self.w_q, self.w_k, self.w_v = attn, w_rpe, norm1, w_q, w_k, w_v

    def forward(self, x, coords, combined_shifts):
        from .attention import attn_front

        q, k, v = attn_front(x, self.norm1, self.w_q, self.w_k, self.w_v, self.attn.num_heads)
        return self.attn(q, k, v, w_rpe=self.w_rpe, coords=coords, combined_shifts=combined_shifts)


def graphed_attn_block(attn: HEPTAttention, w_rpe: nn.Module, norm1: nn.LayerNorm, w_q: nn.Linear, w_k: nn.Linear, w_v: nn.Linear,
                       x, coords, combined_shifts, num_warmup_iters: int = 3):
    """-> callable(x, coords, combined_shifts) replaying captured forward / backward graphs of the front of the reference's
    Attn block (example/transformer.py:157-159) followed by the attention module: two graph launches per step."""
    if not x.is_cuda:
        raise RuntimeError("graphed_attn_block needs CUDA tensors (there is no CPU path)")
    step = _BlockStep(attn, w_rpe, norm1, w_q, w_k, w_v)
    sample = (x.detach().clone().requires_grad_(x.requires_grad), coords.detach().clone(), combined_shifts.detach().clone())
    with torch.cuda.device(x.device):
        return torch.cuda.make_graphed_callables(step, sample, num_warmup_iters=num_warmup_iters, allow_unused_input=True)


class GraphedInference:
    """Forward-only replay of a whole model call (e.g. the pileup Transformer, src/ flavour: BASELINE.json configs[2]) for one
    input shape: ``run(*inputs)`` copies the inputs into the captured buffers, replays, and returns the static output."""

    def __init__(self, model: nn.Module, sample_inputs: Sequence[torch.Tensor], num_warmup_iters: int = 3):
        self.static_in = [t.detach().clone() for t in sample_inputs]
        dev = self.static_in[0].device
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.device(dev), torch.no_grad():
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(num_warmup_iters):
                    model(*self.static_in)
            torch.cuda.current_stream().wait_stream(side)
            with torch.cuda.graph(self.graph):
                self.static_out = model(*self.static_in)

    def run(self, *inputs: torch.Tensor) -> torch.Tensor:
        for dst, src in zip(self.static_in, inputs):
            dst.copy_(src)
        self.graph.replay()
        return self.static_out

"""The caller's LayerNorms on the library's kernels (csrc/layer_norm.cu): the Attn block's ``norm2``
(example/transformer.py:163, src/models/baselines/transformer.py:216) and the 256-wide norms of the model head
(torch_geometric ``MLP(norm="layer_norm")``, example/transformer.py:84).  Same parameters and state_dict keys as
``torch.nn.LayerNorm``; same definition (biased variance, eps inside the square root)."""
import torch
from torch import nn

from . import ops


class _LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, eps: float):
        x2 = x.reshape(-1, x.shape[-1]).contiguous()
        y, mean_rstd = ops.layer_norm_fwd(x2, weight, bias, eps)
        ctx.save_for_backward(x2, mean_rstd, weight)
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        x2, mean_rstd, weight = ctx.saved_tensors
        dx, dw, db = ops.layer_norm_bwd(x2, mean_rstd, weight, dy.reshape(x2.shape).contiguous())
        return dx.view(dy.shape), dw, db, None


class LayerNorm(nn.LayerNorm):
    """``nn.LayerNorm`` over the last dimension; CUDA fp32 inputs of a supported width (4 <= D <= 256, D % 4 == 0) run on the
    library's kernels -- in every mode, so that eval, training and a captured CUDA graph see the same bits; anything else (the
    CPU twin of the tests, other dtypes) runs on torch's."""

    def forward(self, x):
        if (x.is_cuda and x.dtype == torch.float32 and self.elementwise_affine and self.bias is not None
                and len(self.normalized_shape) == 1 and x.numel() > 0 and ops.layer_norm_supported(self.normalized_shape[0])):
            if torch.is_grad_enabled() and (x.requires_grad or self.weight.requires_grad):
                return _LayerNormFn.apply(x, self.weight, self.bias, self.eps)
            y, _ = ops.layer_norm_fwd(x.reshape(-1, x.shape[-1]), self.weight, self.bias, self.eps)   # no autograd node to build
            return y.view(x.shape)
        return super().forward(x)


class _TallLinearFn(torch.autograd.Function):
    """``F.linear`` whose weight gradient is a split-K batched GEMM.  For x (N, in) with N = 60 000 hits and a 256 x 256 (or
    24 x 24) weight, the library's single GEMM dW = dy^T x runs as 16 (or 1) CTAs over the whole contraction -- 0.8 ms per
    256-wide layer of the model head on a B200, 2.4 ms of a 12 ms training step (torch profiler).  Here the hits are cut into S
    slabs, one batched GEMM forms the S partial products and their sum over S is taken in a fixed order (deterministic)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        return torch.nn.functional.linear(x, weight, bias)

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        dy = dy.contiguous()
        dx = dy @ weight if ctx.needs_input_grad[0] else None
        n, tiles = x.shape[0], -(-weight.shape[0] // 64) * -(-weight.shape[1] // 64)
        slabs = max(8, min(256, -(-600 // tiles), n // 256))
        m = n // slabs * slabs
        dw = torch.bmm(dy[:m].view(slabs, m // slabs, -1).transpose(1, 2), x[:m].view(slabs, m // slabs, -1)).sum(0)
        if m < n:
            dw = dw + dy[m:].t() @ x[m:]
        db = None
        if ctx.has_bias:
            if dy.shape[1] >= 64:                     # wide rows: torch's column sum is 3x slower than two stages (72 vs 24 us at 60k x 256)
                mb = n // 240 * 240
                db = dy[:mb].view(240, mb // 240, -1).sum(1).sum(0)
                if mb < n:
                    db = db + dy[mb:].sum(0)
            else:
                db = dy.sum(0)
        return dx, dw, db


class Linear(nn.Linear):
    """``nn.Linear`` (same parameters, same state_dict keys); tall 2-D CUDA inputs (>= 4096 rows) take the split-K weight
    gradient above, everything else torch's own autograd."""

    def forward(self, x):
        if x.is_cuda and x.dim() == 2 and x.shape[0] >= 4096 and torch.is_grad_enabled() and self.weight.requires_grad:
            return _TallLinearFn.apply(x.contiguous(), self.weight, self.bias)
        return super().forward(x)

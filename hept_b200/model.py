"""The caller of the hot path: the HEPT point-cloud Transformer around ``HEPTAttention``.

Mirrors ``Transformer`` / ``Attn`` of the reference with our attention and our batched ``prepare_input``:
  * ``example/transformer.py:66-165``            -> ``flavour="example"`` (batched events, packed integer shifts)
  * ``src/models/baselines/transformer.py:66-229`` (HEPT branch, tasks tracking / pileup) -> ``flavour="src"``
Parameter names and shapes follow the reference so its checkpoints load with ``strict=True``
(``example/ckpt/tracking-60k-model.pt``: ``regions``, ``feat_encoder.*``, ``attns.i.{w_q,w_k,w_v,norm1,norm2,ff,w_rpe}.*``,
``attns.i.attn.{out_linear.*, e2lsh.alpha}``, ``W.weight``, ``mlp_out.lins.j.*``, ``mlp_out.norms.j.*``).

The dense layers are small library GEMMs (cuBLAS through ``nn.Linear``); the attention, the front of the Attn block and the
LayerNorms (``norm2``, the head's norms: ``hept_b200/layers.py``) are the library's CUDA kernels.
``mlp_out`` restates ``torch_geometric.nn.MLP(in, hidden 256, out, num_layers=5, norm="layer_norm", act="tanh")``:
``lin -> LayerNorm -> tanh`` four times, then a plain last ``lin`` (PyG is not a dependency here).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import prepare
from . import ops
from .attention import HEPTAttention, attn_front
from .layers import LayerNorm, Linear


def get_regions(num_regions: int, num_or_hashes: int, num_heads: int, num_and_hashes: int = 2) -> torch.Tensor:
    """eta x phi factorisation of ``num_regions`` per (table, head) -> (T, 2, H); example/hept_utils.py:17-31.

    Random at construction (like the reference); checkpoints overwrite it, parity tests share it via state_dict.
    """
    lb = 2.0
    ub = 2.0 * num_regions ** (1.0 / num_and_hashes) - lb
    raw = torch.rand(num_or_hashes * num_heads, num_and_hashes) * (ub - lb) + lb
    raw = (num_regions / raw.prod(dim=1, keepdim=True)) ** (1.0 / num_and_hashes) * raw
    raw = torch.round(raw * 3) / 3
    return raw.view(num_heads, num_or_hashes, num_and_hashes).permute(1, 2, 0).contiguous()   # "(h c) a -> c a h"


class NodeMLP(nn.Module):
    """PyG ``MLP`` with layer_norm / tanh / plain last layer, same state_dict keys (``lins.j``, ``norms.j``)."""

    def __init__(self, in_channels: int, hidden_channels: int, out_channels: int, num_layers: int):
        super().__init__()
        dims = [in_channels] + [hidden_channels] * (num_layers - 1) + [out_channels]
        self.lins = nn.ModuleList(Linear(a, b) for a, b in zip(dims[:-1], dims[1:]))
        self.norms = nn.ModuleList(LayerNorm(hidden_channels) for _ in range(num_layers - 1))

    def forward(self, x):
        for lin, norm in zip(self.lins[:-1], self.norms):
            x = torch.tanh(norm(lin(x)))
        return self.lins[-1](x)


class Attn(nn.Module):
    """example/transformer.py:131-165; src/models/baselines/transformer.py:160-218 (HEPT branch, pe_type none)."""

    def __init__(self, coords_dim: int, attn_cls=HEPTAttention, **kwargs):
        super().__init__()
        self.dim_per_head = kwargs["h_dim"]
        self.num_heads = kwargs["num_heads"]
        d, h = self.dim_per_head, self.num_heads
        self.w_q = nn.Linear(d, d * h, bias=False)
        self.w_k = nn.Linear(d, d * h, bias=False)
        self.w_v = nn.Linear(d, d * h, bias=False)
        self.attn = attn_cls(d + coords_dim, **kwargs)
        self.dropout = nn.Dropout(0.1)
        self.norm1 = nn.LayerNorm(d)
        self.norm2 = LayerNorm(d)                 # csrc/layer_norm.cu (norm1 is part of the fused front)
        self.ff = nn.Sequential(Linear(d, d), nn.ReLU(), Linear(d, d))
        # eta / phi share the first weight group (example/transformer.py:153-154)
        self.w_rpe = nn.Linear(kwargs["num_w_per_dist"] * (coords_dim - 1), h * d)

    def forward(self, x, kwargs):
        if x.is_cuda and isinstance(self.attn, HEPTAttention) and ops.attn_qkv_supported(self.num_heads, self.dim_per_head):
            q, k, v = attn_front(x, self.norm1, self.w_q, self.w_k, self.w_v, self.num_heads)   # one native call each way
        else:                                   # other shapes, or the oracle-backed twin of the tests: library kernels
            xn = self.norm1(x)
            q, k, v = self.w_q(xn), self.w_k(xn), self.w_v(xn)
        aggr = self.attn(q, k, v, pe=kwargs["coords"], w_rpe=self.w_rpe, **kwargs)
        x = x + self.dropout(aggr)
        return x + self.dropout(self.ff(self.norm2(x)))


class Transformer(nn.Module):
    """HEPT Transformer.  ``flavour="example"``: ``forward(x, coords, batch)``; ``flavour="src"``: one event,
    ``forward(x, coords)`` (``task="pileup"`` adds the particle-id embedding, ``out_proj`` and the sigmoid)."""

    def __init__(self, in_dim: int, coords_dim: int, num_classes: int = 0, dropout: float = 0.1, flavour: str = "example",
                 task: Optional[str] = None, attn_cls=HEPTAttention, prepare_impl=prepare, **kwargs):
        super().__init__()
        assert flavour in ("example", "src")
        self.flavour, self.task = flavour, task
        kwargs.setdefault("e2lsh_beta", flavour == "src")   # src/ checkpoints carry e2lsh.beta (hash_utils.py:344), example/ ones do not
        self._prepare = prepare_impl          # tests swap in an oracle-backed twin (attn_cls + prepare_impl) on CPU
        self.n_layers, self.h_dim = kwargs["n_layers"], kwargs["h_dim"]
        self.num_classes = num_classes
        if task == "pileup":                      # src/models/baselines/transformer.py:76-78
            self.pids_enc = nn.Embedding(7, 10)
            in_dim = in_dim - 1 + 10
        self.feat_encoder = nn.Sequential(Linear(in_dim, self.h_dim), nn.ReLU(), Linear(self.h_dim, self.h_dim))
        self.attns = nn.ModuleList(Attn(coords_dim, attn_cls=attn_cls, **kwargs) for _ in range(self.n_layers))
        self.dropout = nn.Dropout(dropout)
        half = int(self.h_dim // 2)
        self.W = Linear(self.h_dim * (self.n_layers + 1), half, bias=False)
        self.mlp_out = NodeMLP(half, 256, half, num_layers=5)
        self.regions = nn.Parameter(get_regions(kwargs["num_regions"], kwargs["n_hashes"], kwargs["num_heads"]),
                                    requires_grad=False)
        self.helper_params = {"block_size": kwargs["block_size"], "num_heads": kwargs["num_heads"]}
        if task == "pileup":
            self.out_proj = nn.Linear(half, 1)
        elif num_classes:
            self.out_proj = nn.Linear(half, num_classes)

    def forward(self, x, coords, batch=None):
        helper = dict(self.helper_params, regions=self.regions)
        if self.task == "pileup":
            x = torch.cat((x[..., :-1], self.pids_enc(x[..., -1].long())), dim=-1)
        if self.flavour == "example":
            if batch is None:
                batch = torch.zeros(x.shape[0], dtype=torch.long, device=x.device)
            x, kwargs, keep = self._prepare.prepare_input(x, coords, batch, helper)
        else:
            x, kwargs = self._prepare.prepare_input_single(x, coords, helper)
            keep = slice(0, kwargs["raw_size"])
        enc = self.feat_encoder(x)
        stack = [enc]
        for layer in self.attns:
            enc = layer(enc, kwargs)
            stack.append(enc)
        enc = self.W(torch.cat(stack, dim=-1))
        out = enc + self.dropout(self.mlp_out(enc))
        if self.task == "pileup":
            return torch.sigmoid(self.out_proj(out[keep]))
        if self.num_classes:
            out = self.out_proj(out)
        return out[keep]

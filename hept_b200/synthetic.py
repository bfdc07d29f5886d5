"""Synthetic point clouds shaped like the reference's datasets (no network, no 65 GB download).

Shapes follow the reference's dataset transforms:
  tracking: coords = [eta, phi, r, phi_feat, z, eta_rz]  (coords_dim 6, src/datasets/tracking.py:31-32,88)
  pileup:   coords = [eta, phi, f0, f1]                  (coords_dim 4, src/datasets/pileup.py:24,36)

Values are track-like rather than i.i.d. uniform: particles
are drawn in (eta, phi) and each leaves a handful of hits on successive
detector layers with small jitter, so LSH blocks contain genuinely close
points and the kernel-attention numerators do not all underflow.

Everything is generated on CPU from a seeded ``torch.Generator`` so the same
seed gives the same tensors in the build container and on the GPU box.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence

import torch

TRACKING = dict(block_size=100, n_hashes=3, num_regions=150, num_heads=8, h_dim=24, num_w_per_dist=10,
                coords_dim=6, n_layers=4)
PILEUP = dict(block_size=100, n_hashes=3, num_regions=140, num_heads=8, h_dim=24, num_w_per_dist=10,
              coords_dim=4, n_layers=4)

_LAYER_RADII = (0.032, 0.072, 0.116, 0.172, 0.26, 0.36, 0.50, 0.66, 0.82, 1.02)


def point_cloud(n_hits: int, coords_dim: int = 6, seed: int = 0) -> torch.Tensor:
    """(n_hits, coords_dim) float32 coordinates of one event."""
    g = torch.Generator().manual_seed(seed)
    n_part = max(1, n_hits // 12 + 1)
    eta0 = (torch.rand(n_part, generator=g) * 8.0 - 4.0)
    phi0 = (torch.rand(n_part, generator=g) * 2.0 - 1.0) * math.pi
    per = torch.randint(8, 19, (n_part,), generator=g)
    owner = torch.repeat_interleave(torch.arange(n_part), per)
    while owner.numel() < n_hits:                       # top up if the draw came out short
        owner = torch.cat([owner, torch.randint(0, n_part, (n_hits - owner.numel(),), generator=g)])
    owner = owner[torch.randperm(owner.numel(), generator=g)[:n_hits]]
    layer = torch.randint(0, len(_LAYER_RADII), (n_hits,), generator=g)
    r = torch.tensor(_LAYER_RADII)[layer] * (1.0 + 0.01 * torch.randn(n_hits, generator=g))
    eta = eta0[owner] + 0.01 * torch.randn(n_hits, generator=g)
    phi = phi0[owner] + 0.01 * torch.randn(n_hits, generator=g)
    phi = torch.remainder(phi + math.pi, 2 * math.pi) - math.pi
    if coords_dim == 6:
        z = (r * torch.sinh(eta)).clamp(-3.0, 3.0)
        cols = [eta, phi, r, phi / math.pi, z, eta + 0.005 * torch.randn(n_hits, generator=g)]
    elif coords_dim == 4:
        cols = [eta, phi, torch.randn(n_hits, generator=g).abs() * 0.5, torch.rand(n_hits, generator=g)]
    else:
        cols = [eta, phi] + [torch.randn(n_hits, generator=g) for _ in range(coords_dim - 2)]
    return torch.stack(cols, dim=1).float().contiguous()


def module_params(cfg: Dict[str, int], seed: int = 0) -> Dict[str, torch.Tensor]:
    """Default-init parameters with the reference's state_dict names and shapes.

    w_rpe / out_linear follow nn.Linear's default init (kaiming-uniform with
    a=sqrt(5) == U(-1/sqrt(fan_in), 1/sqrt(fan_in))); alpha ~ N(0,1)
    (example/hept_utils.py:42); regions from the reference's recipe
    (example/hept_utils.py:17-31) drawn with our generator.
    """
    g = torch.Generator().manual_seed(seed + 7919)
    h, d, c = cfg["num_heads"], cfg["h_dim"], cfg["coords_dim"]
    t, kk = cfg["n_hashes"], cfg["num_w_per_dist"]
    e = d + c

    def lin(out_f, in_f):
        b = 1.0 / math.sqrt(in_f)
        return (torch.rand(out_f, in_f, generator=g) * 2 - 1) * b, (torch.rand(out_f, generator=g) * 2 - 1) * b

    w_rpe_w, w_rpe_b = lin(h * d, kk * (c - 1))
    out_w, out_b = lin(d, h * d)
    alpha = torch.randn(h, e, t, generator=g)
    lb, nr = 2.0, float(cfg["num_regions"])
    ub = 2.0 * nr ** 0.5 - lb
    raw = torch.rand(t * h, 2, generator=g) * (ub - lb) + lb
    raw = (nr / raw.prod(dim=1, keepdim=True)) ** 0.5 * raw
    raw = torch.round(raw * 3) / 3
    regions = raw.view(h, t, 2).permute(1, 2, 0).contiguous()      # "(h c) a -> c a h"
    return {
        "w_rpe.weight": w_rpe_w.contiguous(), "w_rpe.bias": w_rpe_b.contiguous(),
        "out_linear.weight": out_w.contiguous(), "out_linear.bias": out_b.contiguous(),
        "e2lsh.alpha": alpha.contiguous(), "regions": regions,
    }


def qkv(n: int, cfg: Dict[str, int], seed: int = 0, std: float = 0.5):
    g = torch.Generator().manual_seed(seed + 104729)
    width = cfg["num_heads"] * cfg["h_dim"]
    return tuple((torch.randn(n, width, generator=g) * std).contiguous() for _ in range(3))


def tracking_truth(n_hits: int, seed: int = 0, random_pairs_per_hit: int = 32):
    """Synthetic truth of one tracking event for the loss / metrics (src/tracking_trainer.py:26, src/utils/losses.py):
    ``cluster_ids`` (64-bit particle ids, ~12 hits each, 3 % noise hits with id 0), ``recons`` (0 / 1), ``pts``, and
    ``point_pairs`` (2, P): every hit paired with its particle mates and with ``random_pairs_per_hit`` other hits (the
    dataset ships radius-graph pairs capped at 256 per hit, src/datasets/tracking.py:153), shuffled."""
    g = torch.Generator().manual_seed(seed + 31337)
    n_part = max(1, n_hits // 12)
    owner = torch.randint(0, n_part, (n_hits,), generator=g)
    cid = (owner.long() + 1) * 4503599627370497 % (1 << 62)
    cid[torch.rand(n_hits, generator=g) < 0.03] = 0
    recons = (torch.rand(n_hits, generator=g) < 0.9).float()
    pts = torch.rand(n_hits, generator=g) * 3.0
    src = torch.arange(n_hits).repeat_interleave(random_pairs_per_hit)
    dst = torch.randint(0, n_hits, (n_hits * random_pairs_per_hit,), generator=g)
    order = torch.argsort(owner, stable=True)
    so = owner[order]
    mates = []
    for shift in range(1, 12):                       # hits of one particle are adjacent in `order`
        same = so[shift:] == so[:-shift]
        a, b = order[shift:][same], order[:-shift][same]
        mates.append(torch.stack([a, b]))
        mates.append(torch.stack([b, a]))
    pairs = torch.cat([torch.stack([src, dst])] + mates, dim=1)
    pairs = pairs[:, pairs[0] != pairs[1]]
    pairs = pairs[:, torch.randperm(pairs.shape[1], generator=g)].contiguous()
    return cid, recons, pts, pairs


def event_sizes(kind: str) -> List[int]:
    """Raw hit counts of the BASELINE.json configs."""
    return {
        "tracking-6k": [6037],
        "tracking-60k": [60000],
        "tracking-60k-ragged": [61237],
        "pileup-10k": [10000],
        "batched-imbalanced": [21000, 15000, 11300, 9000, 3000, 700, 130, 57],
    }[kind]


def batched_cloud(sizes: Sequence[int], coords_dim: int = 6, seed: int = 0):
    """Concatenated events + ascending ``batch`` vector, as the example/ path expects."""
    coords = torch.cat([point_cloud(s, coords_dim, seed + 31 * i) for i, s in enumerate(sizes)])
    batch = torch.repeat_interleave(torch.arange(len(sizes)), torch.tensor(list(sizes)))
    return coords, batch

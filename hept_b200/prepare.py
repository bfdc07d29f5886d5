"""Per-forward preparation of the AND-hash inputs (rows a13-a17 of SURVEY.md 8(a)).

Host-side mirror of the reference's ``prepare_input``:
  * ``prepare_input(x, coords, batch, helper_params)``  — example/transformer.py:35-63 (batched events):
    per-event eta/phi rank -> quantile region (example/hept_utils.py:6-14) -> bit-pack phi over eta and the
    batch index over both (example/transformer.py:10-13) -> pad every event to a block multiple by
    repeating real points (example/transformer.py:16-32);
  * ``prepare_input_single(x, coords, helper_funcs)``   — HEPT branch of
    src/models/baselines/transformer.py:43-57 (one event: zero / +inf padding, float region indices).

Unlike the reference there is no Python loop over events: ranks inside events come from two stable
sorts over the whole batch, and the padding plan is built with segment arithmetic, so the work is a
fixed handful of device launches whatever the number of events.  Runs on whatever device the inputs
are on (round 1: torch ops; the same results as the reference, bit for bit, except where the
reference's unstable argsort breaks ties differently — see ``tests/test_prepare.py``).
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch


def _rank_in_event(values: torch.Tensor, batch: torch.Tensor, starts: torch.Tensor) -> torch.Tensor:
    """rank of every point among the points of its own event, by ascending ``values`` (stable)."""
    by_val = torch.argsort(values, stable=True)
    by_evt = torch.argsort(batch[by_val], stable=True)
    order = by_val[by_evt]                                    # sorted by (event, value)
    rank = torch.empty_like(order)
    rank[order] = torch.arange(order.numel(), device=order.device) - starts[batch[order]]
    return rank


def _quantile_region(rank: torch.Tensor, event_size: torch.Tensor, num_regions: torch.Tensor) -> torch.Tensor:
    """floor(rank / ceil(n_event / num_regions)) + 1 in the reference's dtypes -> (T*H, N) float32.

    example/hept_utils.py:8-12: ``region_size = ceil(n / num_regions)`` (float32), then an int64 arange is
    floor-divided by it (float32 result).
    """
    # ``n / num_regions`` with a Python int on the left is Tensor.__rtruediv__ == reciprocal() * n: two
    # roundings, and ceil() sees the difference (700 / 14.0 -> 50.000004 -> 51)
    width = torch.ceil(torch.reciprocal(num_regions)[:, None] * event_size[None, :].to(num_regions.dtype))  # (TH, N)
    return rank[None, :] // width + 1


def _pack(low: torch.Tensor, high: torch.Tensor) -> torch.Tensor:
    """example/transformer.py:10-13 — (high << ceil(log2(max(low)+1))) | low, width chosen per row."""
    top = low.max(dim=1, keepdim=True).values
    bits = torch.ceil(torch.log2(top + 1)).long()
    return (high << bits) | low


def prepare_input(x: torch.Tensor, coords: torch.Tensor, batch: torch.Tensor, helper_params: Dict):
    """-> (x_padded, {"combined_shifts": (T,H,Np) int64, "coords": (Np,C)}, unpad_mask (Np,) bool)."""
    regions = helper_params["regions"]                        # (T, 2, H)
    block, heads = int(helper_params["block_size"]), int(helper_params["num_heads"])
    t = regions.shape[0]
    reg = regions.permute(1, 0, 2).reshape(2, t * heads)       # "c a h -> a (c h)"
    with torch.no_grad():
        dev = coords.device
        batch = batch.long()
        sizes = torch.bincount(batch)
        ends = sizes.cumsum(0)
        starts = ends - sizes
        size_of = sizes[batch]
        eta = _quantile_region(_rank_in_event(coords[:, 0], batch, starts), size_of, reg[0]).long()
        phi = _quantile_region(_rank_in_event(coords[:, 1], batch, starts), size_of, reg[1]).long()
        code = _pack(_pack(eta, phi), batch[None]).view(t, heads, -1)

        # padding plan: event i gets pad_i extra rows copied from order[ends[i] - block + j], j < pad_i,
        # where ``order`` sorts all raw points by the (table 0, head 0) code (example/transformer.py:23-30)
        padded = (sizes + block - 1) // block * block
        pads = padded - sizes
        n_raw, n_pad = int(sizes.sum()), int(padded.sum())
        pad_starts = padded.cumsum(0) - padded
        order = torch.argsort(code[0, 0], stable=True)
        evt = torch.repeat_interleave(torch.arange(len(sizes), device=dev), padded, output_size=n_pad)
        local = torch.arange(n_pad, device=dev) - pad_starts[evt]
        is_real = local < sizes[evt]
        src = ends[evt] - block + (local - sizes[evt])          # only meaningful on pad rows
        src = torch.where(src < 0, src + n_raw, src)            # Python-style wrap, as tensor indexing does
        take = torch.where(is_real, starts[evt] + local, order[src.clamp(0, n_raw - 1)])
        kwargs = {"combined_shifts": code[..., take].contiguous(), "coords": coords[take]}
    return x[take], kwargs, is_real


def prepare_input_single(x: torch.Tensor, coords: torch.Tensor, helper_funcs: Dict):
    """-> (x_padded, {"coords", "raw_size", "regions_h", "region_indices": [eta, phi]})."""
    regions = helper_funcs["regions"]
    block = int(helper_funcs["block_size"])
    t, _, heads = regions.shape
    with torch.no_grad():
        n = x.shape[0]
        pad = (-n) % block
        regions_h = regions.permute(1, 0, 2).reshape(2, t * heads)
        if pad:
            x = torch.cat([x, x.new_zeros((pad,) + tuple(x.shape[1:]))])
            coords = torch.cat([coords, coords.new_full((pad, coords.shape[1]), float("inf"))])
        else:
            coords = coords.clone()
        total = coords.shape[0]
        zero = torch.zeros(total, dtype=torch.long, device=coords.device)
        size_of = torch.full((total,), total, device=coords.device)
        eta = _quantile_region(_rank_in_event(coords[:, 0], zero, zero[:1]), size_of, regions_h[0])
        phi = _quantile_region(_rank_in_event(coords[:, 1], zero, zero[:1]), size_of, regions_h[1])
        coords[n:] = 0.0
    return x, {"coords": coords, "raw_size": n, "regions_h": regions_h, "region_indices": [eta, phi]}

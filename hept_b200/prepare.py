"""Per-forward preparation of the AND-hash inputs (rows a13-a17 of SURVEY.md 8(a)) on the library's CUDA kernels.

Host-side mirror of the reference's ``prepare_input``:
  * ``prepare_input(x, coords, batch, helper_params)``  — example/transformer.py:35-63 (batched events):
    per-event eta/phi rank -> quantile region (example/hept_utils.py:6-14) -> bit-pack phi over eta and the
    batch index over both (example/transformer.py:10-13) -> pad every event to a block multiple by
    repeating real points (example/transformer.py:16-32);
  * ``prepare_input_single(x, coords, helper_funcs)``   — HEPT branch of
    src/models/baselines/transformer.py:43-57 (one event: zero / +inf padding, float region indices).

Everything between the raw coordinates and the padded codes runs in libhept_sm100.so (``hept_prepare_batched`` /
``hept_prepare_single``, csrc/prepare.cu): one segmented argsort for the eta / phi ranks of all events, one for the
padding order, five small kernels.  There is no per-event Python loop and no CPU path.  The only host work is what decides
tensor SIZES: the event sizes (``batch.bincount()`` read back once, as the reference does; pass ``sizes=`` to skip even
that) and the prefix sums of those few integers.  Results equal the reference's bit for bit except where its unstable
argsort breaks ties between EQUAL eta / phi / code values differently (tests/test_prepare.py).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence

import torch

from . import ops


_OFFSET_CACHE: Dict = {}


def _regions_h(regions: torch.Tensor):
    """(T, 2, H) -> ((2, T*H) device tensor: the reference's rearrange "c a h -> a (c h)" (example/transformer.py:37), and
    the bits one region index needs).  ``regions`` is a frozen parameter: both are computed once per tensor version and kept
    ON the tensor object (a cache keyed by address would outlive the tensor and hit a stranger at the same address)."""
    hit = getattr(regions, "_hept_regions_h", None)
    if hit is None or hit[0] != regions._version:
        t, two, heads = regions.shape
        rh = regions.detach().permute(1, 0, 2).reshape(two, t * heads).contiguous().float()
        top = int(math.ceil(float(rh.max()))) + 2           # region = floor(rank / ceil(n / r)) + 1 <= ceil(r) + 1
        hit = (regions._version, rh, max(1, (top - 1).bit_length()))
        try:
            regions._hept_regions_h = hit
        except AttributeError:
            pass
    return hit[1], hit[2]


def _offsets(sizes, block: int, dev):
    """int32 device tensor [event_start (E+1) | pad_start (E+1)] + (n_raw, n_pad); cached per (sizes, block, device)."""
    key = (tuple(sizes), block, dev)
    hit = _OFFSET_CACHE.get(key)
    if hit is None:
        ev_start, pad_start = [0], [0]
        for s in sizes:
            ev_start.append(ev_start[-1] + s)
            pad_start.append(pad_start[-1] + (s + block - 1) // block * block)
        if len(_OFFSET_CACHE) > 256:
            _OFFSET_CACHE.clear()
        hit = _OFFSET_CACHE[key] = (torch.tensor(ev_start + pad_start, dtype=torch.int32, device=dev), ev_start[-1], pad_start[-1])
    return hit


def prepare_input(x: torch.Tensor, coords: torch.Tensor, batch: torch.Tensor, helper_params: Dict,
                  sizes: Optional[Sequence[int]] = None):
    """-> (x_padded, {"combined_shifts": (T,H,Np) int64, "combined_shifts32": int32 copy, "coords": (Np,C)}, unpad (Np,) bool).

    ``batch`` must be ascending (the reference slices events by cumulative size, example/transformer.py:44-48)."""
    if not coords.is_cuda:
        raise RuntimeError("hept_b200.prepare_input runs on CUDA (sm_100a) only; there is no CPU path")
    regions = helper_params["regions"]
    block, heads = int(helper_params["block_size"]), int(helper_params["num_heads"])
    t = regions.shape[0]
    with torch.no_grad():
        dev = coords.device
        batch = batch.to(device=dev, dtype=torch.int64)
        if sizes is None:
            sizes = torch.bincount(batch).tolist()          # the one host read-back: it decides the output sizes
        sizes = [int(s) for s in sizes]
        offsets, n_raw, n_pad = _offsets(sizes, block, dev)
        if n_raw != coords.shape[0]:
            raise ValueError(f"event sizes sum to {n_raw} but there are {coords.shape[0]} points")
        regions_h, region_bits = _regions_h(regions if regions.device == dev else regions.to(dev))
        code_bits = max(1, (len(sizes) - 1).bit_length()) + 2 * region_bits      # batch index over phi over eta
        shifts, shifts32, take, real, coords_pad = ops.prepare_batched(
            coords.float(), batch, offsets, len(sizes), n_raw, n_pad, max(sizes), regions_h, block,
            code_bits=code_bits if code_bits <= 32 else 0)
        kwargs = {"combined_shifts": shifts.view(t, heads, n_pad), "combined_shifts32": shifts32.view(t, heads, n_pad),
                  "coords": coords_pad}
    return x[take], kwargs, real


def prepare_input_single(x: torch.Tensor, coords: torch.Tensor, helper_funcs: Dict):
    """-> (x_padded, {"coords", "raw_size", "regions_h", "region_indices": [eta, phi]})."""
    if not coords.is_cuda:
        raise RuntimeError("hept_b200.prepare_input_single runs on CUDA (sm_100a) only; there is no CPU path")
    regions = helper_funcs["regions"]
    block = int(helper_funcs["block_size"])
    with torch.no_grad():
        n = x.shape[0]
        pad = (-n) % block
        regions_h, _ = _regions_h(regions if regions.device == coords.device else regions.to(coords.device))
        if pad:
            x = torch.cat([x, x.new_zeros((pad,) + tuple(x.shape[1:]))])
        coords_pad, eta, phi = ops.prepare_single(coords.float(), n + pad, regions_h)
    return x, {"coords": coords_pad, "raw_size": n, "regions_h": regions_h, "region_indices": [eta, phi]}

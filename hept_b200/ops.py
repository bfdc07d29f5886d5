"""Stage-wise torch wrappers over the C ABI (include/hept_b200.h).

torch is plumbing here: it owns device memory and the current CUDA stream; every function below
validates its tensors, allocates outputs / workspace from torch's caching allocator and hands raw
device pointers to libhept_sm100.so.  Nothing here computes on the host and nothing falls back.
The stage-wise entry points exist so parity tests can inject the oracle's intermediates
(SURVEY.md 8(b)).
"""
from __future__ import annotations

import ctypes as C
import functools
from dataclasses import dataclass
from typing import Optional, Sequence, Tuple

import torch

from . import _lib

STAGE_ROW = 32  # floats per staged (hit, head, table) row


@dataclass(frozen=True)
class Dims:
    N: int
    H: int
    D: int
    C: int
    T: int
    B: int
    raw_size: int

    def struct(self) -> _lib.Shape:
        return _lib.Shape(self.N, self.H, self.D, self.C, self.T, self.B, self.raw_size)

    @property
    def E(self) -> int:
        return self.D + self.C


def _stream(t: torch.Tensor) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _tensors(args):
    for a in args:
        if isinstance(a, torch.Tensor):
            yield a
        elif isinstance(a, (list, tuple)):
            yield from _tensors(a)


def _on_device(fn):
    """Run a native stage with the tensors' device current.

    The library launches on the CURRENT CUDA device and keeps per-device state (include/hept_b200.h): every tensor
    argument must live on one CUDA device, and that device is made current for the call, so a module placed on
    ``cuda:1`` with ``.to()`` works without the caller ever calling ``torch.cuda.set_device``.
    """

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        dev = None
        for t in _tensors(list(args) + list(kwargs.values())):
            if not t.is_cuda:
                raise RuntimeError(f"{fn.__name__}: tensor on {t.device}: hept_b200 runs on CUDA (sm_100a) only, "
                                   "there is no CPU path")
            if dev is None:
                dev = t.device
            elif t.device != dev:
                raise RuntimeError(f"{fn.__name__}: tensors on different devices ({dev} and {t.device})")
        if dev is None:
            return fn(*args, **kwargs)
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)

    return wrapper


def _ptr(t: Optional[torch.Tensor]) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def _need(t: torch.Tensor, name: str, dtype, shape: Optional[Sequence[int]] = None) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{name} is on {t.device}: hept_b200 runs on CUDA (sm_100a) only, there is no CPU path")
    if t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise ValueError(f"{name} must have shape {tuple(shape)}, got {tuple(t.shape)}")
    return t if t.is_contiguous() else t.contiguous()


def _workspace(nbytes: int, like: torch.Tensor) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=like.device)


def supported(D: int, C_: int, B: int) -> bool:
    return bool(_lib.load().hept_shape_supported(D, C_, B))


def launch_count(reset: bool = False) -> int:
    return int(_lib.load().hept_launch_count(1 if reset else 0))


# ------------------------------------------------------------------------------------------ stages
@_on_device
def coord_scale(w_rpe_weight: torch.Tensor, H: int, D: int, K: int) -> torch.Tensor:
    lib = _lib.load()
    w = _need(w_rpe_weight, "w_rpe.weight", torch.float32)
    if w.dim() != 2 or w.shape[0] != H * D or w.shape[1] % K:
        raise ValueError(f"w_rpe.weight must be (H*D={H * D}, R*K) with K={K}, got {tuple(w.shape)}")
    R = w.shape[1] // K
    scale = torch.empty(H, R + 1, dtype=torch.float32, device=w.device)
    _lib.check(lib.hept_coord_scale_fwd(_ptr(w), H, D, R, K, _ptr(scale), _stream(w)), "hept_coord_scale_fwd")
    return scale


@_on_device
def coord_scale_backward(w_rpe_weight: torch.Tensor, scale: torch.Tensor, dscale: torch.Tensor, H: int, D: int,
                         K: int) -> torch.Tensor:
    lib = _lib.load()
    w = _need(w_rpe_weight, "w_rpe.weight", torch.float32)
    R = w.shape[1] // K
    scale = _need(scale, "scale", torch.float32, (H, R + 1))
    dscale = _need(dscale, "dscale", torch.float32, (H, R + 1))
    dw = torch.empty_like(w)
    _lib.check(lib.hept_coord_scale_bwd(_ptr(w), _ptr(scale), _ptr(dscale), H, D, R, K, _ptr(dw), _stream(w)),
               "hept_coord_scale_bwd")
    return dw


@_on_device
def hash_project(d: Dims, q, k, coords, scale, alpha) -> Tuple[torch.Tensor, torch.Tensor]:
    """-> proj (2, T, H, N), span (T, H)."""
    lib = _lib.load()
    q = _need(q, "query", torch.float32, (d.N, d.H * d.D))
    k = _need(k, "key", torch.float32, (d.N, d.H * d.D))
    coords = _need(coords, "coords", torch.float32, (d.N, d.C))
    scale = _need(scale, "scale", torch.float32, (d.H, d.C))
    alpha = _need(alpha, "e2lsh.alpha", torch.float32, (d.H, d.E, d.T))
    proj = torch.empty(2, d.T, d.H, d.N, dtype=torch.float32, device=q.device)
    span = torch.empty(d.T, d.H, dtype=torch.float32, device=q.device)
    s = d.struct()
    ws = _workspace(lib.hept_hash_workspace_bytes(C.byref(s)), q)
    _lib.check(lib.hept_hash_project(C.byref(s), _ptr(q), _ptr(k), _ptr(coords), _ptr(scale), _ptr(alpha), _ptr(proj),
                                     _ptr(span), _ptr(ws), ws.numel(), _stream(q)), "hept_hash_project")
    return proj, span


@_on_device
def keys_from_packed_shifts(d: Dims, proj, span, combined_shifts) -> torch.Tensor:
    lib = _lib.load()
    proj = _need(proj, "proj", torch.float32, (2, d.T, d.H, d.N))
    span = _need(span, "span", torch.float32, (d.T, d.H))
    sh32 = combined_shifts.dtype == torch.int32
    sh = _need(combined_shifts, "combined_shifts", torch.int32 if sh32 else torch.int64, (d.T, d.H, d.N))
    keys = torch.empty_like(proj)
    s = d.struct()
    fn = lib.hept_keys_from_packed_shifts32 if sh32 else lib.hept_keys_from_packed_shifts
    _lib.check(fn(C.byref(s), _ptr(proj), _ptr(span), _ptr(sh), _ptr(keys), _stream(proj)), "hept_keys_from_packed_shifts")
    return keys


@_on_device
def keys_from_region_indices(d: Dims, proj, span, region_eta, region_phi, regions_h) -> torch.Tensor:
    lib = _lib.load()
    proj = _need(proj, "proj", torch.float32, (2, d.T, d.H, d.N))
    span = _need(span, "span", torch.float32, (d.T, d.H))
    eta = _need(region_eta, "region_indices[0]", torch.float32, (d.T * d.H, d.N))
    phi = _need(region_phi, "region_indices[1]", torch.float32, (d.T * d.H, d.N))
    rh = _need(regions_h, "regions_h", torch.float32, (2, d.T * d.H))
    keys = torch.empty_like(proj)
    s = d.struct()
    _lib.check(lib.hept_keys_from_region_indices(C.byref(s), _ptr(proj), _ptr(span), _ptr(eta), _ptr(phi), _ptr(rh),
                                                 _ptr(keys), _stream(proj)), "hept_keys_from_region_indices")
    return keys


@_on_device
def segmented_argsort(keys: torch.Tensor) -> torch.Tensor:
    """Stable ascending argsort along the last dim of a float32 tensor -> int32 positions, same shape."""
    lib = _lib.load()
    keys = _need(keys, "keys", torch.float32)
    n = keys.shape[-1]
    segs = keys.numel() // n
    pos = torch.empty(keys.shape, dtype=torch.int32, device=keys.device)
    nbytes = lib.hept_argsort_workspace_bytes(segs, n)
    ws = _workspace(nbytes, keys)
    _lib.check(lib.hept_segmented_argsort(_ptr(keys), segs, n, _ptr(pos), _ptr(ws), ws.numel(), _stream(keys)),
               "hept_segmented_argsort")
    return pos


@_on_device
def hat_coords(d: Dims, coords, scale) -> torch.Tensor:
    """-> (N, H, 8): scale[h,c] * coords[n,c], zero padded (the coordinate part of q_hat / k_hat)."""
    lib = _lib.load()
    coords = _need(coords, "coords", torch.float32, (d.N, d.C))
    scale = _need(scale, "scale", torch.float32, (d.H, d.C))
    hat = torch.empty(d.N, d.H, 8, dtype=torch.float32, device=coords.device)
    s = d.struct()
    _lib.check(lib.hept_hat_coords(C.byref(s), _ptr(coords), _ptr(scale), _ptr(hat), _stream(coords)), "hept_hat_coords")
    return hat


@_on_device
def block_attention_fwd(d: Dims, q, k, v, coords, scale, positions) -> torch.Tensor:
    """-> stage (H, N, T, 32): numerator [0:D) and normaliser [D] per (head, hit, table), original hit order."""
    lib = _lib.load()
    q = _need(q, "query", torch.float32, (d.N, d.H * d.D))
    k = _need(k, "key", torch.float32, (d.N, d.H * d.D))
    v = _need(v, "value", torch.float32, (d.N, d.H * d.D))
    coords = _need(coords, "coords", torch.float32, (d.N, d.C))
    scale = _need(scale, "scale", torch.float32, (d.H, d.C))
    pos = _need(positions, "positions", torch.int32, (2, d.T, d.H, d.N))
    stage = torch.empty(d.H, d.N, d.T, STAGE_ROW, dtype=torch.float32, device=q.device)
    hat = hat_coords(d, coords, scale)
    s = d.struct()
    _lib.check(lib.hept_block_attention_fwd(C.byref(s), _ptr(q), _ptr(k), _ptr(v), _ptr(coords), _ptr(scale), _ptr(hat),
                                            _ptr(pos), _ptr(stage), _stream(q)), "hept_block_attention_fwd")
    return stage


@_on_device
def or_combine(d: Dims, stage) -> Tuple[torch.Tensor, torch.Tensor]:
    lib = _lib.load()
    stage = _need(stage, "stage", torch.float32, (d.H, d.N, d.T, STAGE_ROW))
    out_pre = torch.empty(d.N, d.H * d.D, dtype=torch.float32, device=stage.device)
    den = torch.empty(d.N, d.H, dtype=torch.float32, device=stage.device)
    s = d.struct()
    _lib.check(lib.hept_or_combine(C.byref(s), _ptr(stage), _ptr(out_pre), _ptr(den), _stream(stage)), "hept_or_combine")
    return out_pre, den


@_on_device
def out_linear_fwd(d: Dims, out_pre, weight, bias) -> torch.Tensor:
    """out (N, D) = out_pre (N, H*D) weight^T + bias   (example/hept.py:80)."""
    lib = _lib.load()
    x = _need(out_pre, "out_pre", torch.float32, (d.N, d.H * d.D))
    w = _need(weight, "out_linear.weight", torch.float32, (d.D, d.H * d.D))
    b = _need(bias, "out_linear.bias", torch.float32, (d.D,))
    out = torch.empty(d.N, d.D, dtype=torch.float32, device=x.device)
    s = d.struct()
    _lib.check(lib.hept_out_linear_fwd(C.byref(s), _ptr(x), _ptr(w), _ptr(b), _ptr(out), _stream(x)), "hept_out_linear_fwd")
    return out


@_on_device
def out_linear_bwd(d: Dims, d_out, weight, out_pre, need_input_grad: bool = True):
    """-> d_out_pre (N, H*D) or None, d_weight (D, H*D), d_bias (D)."""
    lib = _lib.load()
    g = _need(d_out, "d_out", torch.float32, (d.N, d.D))
    w = _need(weight, "out_linear.weight", torch.float32, (d.D, d.H * d.D))
    x = _need(out_pre, "out_pre", torch.float32, (d.N, d.H * d.D))
    dx = torch.empty_like(x) if need_input_grad else None
    dw = torch.empty_like(w)
    db = torch.empty(d.D, dtype=torch.float32, device=x.device)
    s = d.struct()
    ws = _workspace(lib.hept_out_linear_bwd_workspace_bytes(C.byref(s)), x)
    _lib.check(lib.hept_out_linear_bwd(C.byref(s), _ptr(g), _ptr(w), _ptr(x), _ptr(dx), _ptr(dw), _ptr(db), _ptr(ws),
                                       ws.numel(), _stream(x)), "hept_out_linear_bwd")
    return dx, dw, db


# ---------------------------------------------------------------- the caller's other LayerNorms (norm2, the model head)
def layer_norm_supported(D: int) -> bool:
    return bool(_lib.load().hept_layer_norm_supported(D))


@_on_device
def layer_norm_fwd(x, weight, bias, eps: float):
    """x (N, D) -> y (N, D), mean_rstd (N, 2) (kept for the backward)."""
    lib = _lib.load()
    n, d = x.shape
    x = _need(x, "x", torch.float32, (n, d))
    g = _need(weight, "weight", torch.float32, (d,))
    b = _need(bias, "bias", torch.float32, (d,))
    y = torch.empty_like(x)
    mr = torch.empty(n, 2, dtype=torch.float32, device=x.device)
    _lib.check(lib.hept_layer_norm_fwd(_ptr(x), _ptr(g), _ptr(b), n, d, eps, _ptr(y), _ptr(mr), _stream(x)), "hept_layer_norm_fwd")
    return y, mr


@_on_device
def layer_norm_bwd(x, mean_rstd, weight, dy):
    """-> dx (N, D), d weight, d bias (D)."""
    lib = _lib.load()
    n, d = x.shape
    x = _need(x, "x", torch.float32, (n, d))
    mr = _need(mean_rstd, "mean_rstd", torch.float32, (n, 2))
    g = _need(weight, "weight", torch.float32, (d,))
    dy = _need(dy, "dy", torch.float32, (n, d))
    dx = torch.empty_like(x)
    dg, db = torch.empty(d, dtype=torch.float32, device=x.device), torch.empty(d, dtype=torch.float32, device=x.device)
    ws = _workspace(lib.hept_layer_norm_bwd_workspace_bytes(n, d), x)
    _lib.check(lib.hept_layer_norm_bwd(_ptr(x), _ptr(mr), _ptr(g), _ptr(dy), n, d, _ptr(dx), _ptr(dg), _ptr(db), _ptr(ws), ws.numel(),
                                       _stream(x)), "hept_layer_norm_bwd")
    return dx, dg, db


# ---------------------------------------------------------------- norm1 + w_q / w_k / w_v
def attn_qkv_supported(H: int, D: int) -> bool:
    return bool(_lib.load().hept_attn_qkv_supported(H, D))


@_on_device
def attn_qkv_fwd(x, norm_weight, norm_bias, w_q, w_k, w_v, H: int, D: int, eps: float):
    """x (N, D) -> q, k, v (N, H*D) = w(norm1(x)); also returns x_normed (N, D) and the transposed weights wt (3, D, H*D)."""
    lib = _lib.load()
    n = x.shape[0]
    x = _need(x, "x", torch.float32, (n, D))
    g = _need(norm_weight, "norm1.weight", torch.float32, (D,))
    b = _need(norm_bias, "norm1.bias", torch.float32, (D,))
    ws_ = [_need(w, nm, torch.float32, (H * D, D)) for w, nm in ((w_q, "w_q.weight"), (w_k, "w_k.weight"), (w_v, "w_v.weight"))]
    dev = x.device
    wt = torch.empty(3, D, H * D, dtype=torch.float32, device=dev)
    xn = torch.empty(n, D, dtype=torch.float32, device=dev)
    q, k, v = (torch.empty(n, H * D, dtype=torch.float32, device=dev) for _ in range(3))
    _lib.check(lib.hept_attn_qkv_fwd(_ptr(x), _ptr(g), _ptr(b), _ptr(ws_[0]), _ptr(ws_[1]), _ptr(ws_[2]), n, H, D, float(eps),
                                     _ptr(wt), _ptr(xn), _ptr(q), _ptr(k), _ptr(v), _stream(x)), "hept_attn_qkv_fwd")
    return q, k, v, xn, wt


@_on_device
def attn_qkv_bwd(x, xn, norm_weight, wt, dq, dk, dv, H: int, D: int, eps: float):
    """-> dx (N, D), d norm1.weight, d norm1.bias (D), d w_q, d w_k, d w_v (H*D, D)."""
    lib = _lib.load()
    n = x.shape[0]
    x = _need(x, "x", torch.float32, (n, D))
    xn = _need(xn, "x_normed", torch.float32, (n, D))
    g = _need(norm_weight, "norm1.weight", torch.float32, (D,))
    wt = _need(wt, "wt", torch.float32, (3, D, H * D))
    dq, dk, dv = (_need(t, nm, torch.float32, (n, H * D)) for t, nm in ((dq, "dq"), (dk, "dk"), (dv, "dv")))
    dev = x.device
    dx = torch.empty(n, D, dtype=torch.float32, device=dev)
    dgam, dbet = torch.empty(D, dtype=torch.float32, device=dev), torch.empty(D, dtype=torch.float32, device=dev)
    dwq, dwk, dwv = (torch.empty(H * D, D, dtype=torch.float32, device=dev) for _ in range(3))
    ws = _workspace(lib.hept_attn_qkv_bwd_workspace_bytes(n, H, D), x)
    _lib.check(lib.hept_attn_qkv_bwd(_ptr(x), _ptr(xn), _ptr(g), _ptr(wt), _ptr(dq), _ptr(dk), _ptr(dv), n, H, D, float(eps),
                                     _ptr(dx), _ptr(dgam), _ptr(dbet), _ptr(dwq), _ptr(dwk), _ptr(dwv), _ptr(ws), ws.numel(),
                                     _stream(x)), "hept_attn_qkv_bwd")
    return dx, dgam, dbet, dwq, dwk, dwv


# ------------------------------------------------------------------ InfoNCE loss
METRICS = {"l2_rbf": 0, "l2_inverse": 1, "cosine": 2}


@_on_device
def infonce_fwd(x, point_pairs, cluster_ids, recons, pts, metric: str, tau: float, pt_thres: float = 0.9):
    """-> loss (0-dim tensor), saved state (uint8 tensor) for infonce_bwd."""
    lib = _lib.load()
    x = _need(x, "x", torch.float32)
    n, d = x.shape
    pairs = _need(point_pairs, "point_pairs", torch.int64)
    if pairs.dim() != 2 or pairs.shape[0] != 2:
        raise ValueError(f"point_pairs must be (2, P), got {tuple(pairs.shape)}")
    p = pairs.shape[1]
    cid = _need(cluster_ids, "cluster_ids", torch.int64, (n,))
    recons = _need(recons, "recons", torch.float32, (n,))
    pts = _need(pts, "pts", torch.float32, (n,))
    loss = torch.empty((), dtype=torch.float32, device=x.device)
    saved = _workspace(lib.hept_infonce_saved_bytes(n, p), x)
    ws = _workspace(lib.hept_infonce_workspace_bytes(n, p, 0), x)
    _lib.check(lib.hept_infonce_fwd(_ptr(x), n, d, _ptr(pairs), p, _ptr(cid), _ptr(recons), _ptr(pts), float(pt_thres),
                                    METRICS[metric], float(tau), _ptr(loss), _ptr(saved), saved.numel(), _ptr(ws), ws.numel(),
                                    _stream(x)), "hept_infonce_fwd")
    return loss, saved


@_on_device
def infonce_bwd(x, point_pairs, saved, grad_loss, metric: str, tau: float):
    lib = _lib.load()
    x = _need(x, "x", torch.float32)
    n, d = x.shape
    pairs = _need(point_pairs, "point_pairs", torch.int64)
    p = pairs.shape[1]
    g = _need(grad_loss.reshape(1), "grad_loss", torch.float32, (1,))
    dx = torch.empty_like(x)
    ws = _workspace(lib.hept_infonce_workspace_bytes(n, p, 1), x)
    _lib.check(lib.hept_infonce_bwd(_ptr(x), n, d, _ptr(pairs), p, METRICS[metric], float(tau), _ptr(g), _ptr(saved), saved.numel(),
                                    _ptr(dx), _ptr(ws), ws.numel(), _stream(x)), "hept_infonce_bwd")
    return dx


@_on_device
def knn_metrics(x, cluster_ids, queries, cosine: bool, K: int):
    """-> 5 floats on the device: mean accuracy, precision, recall, number of scored queries, largest k seen."""
    lib = _lib.load()
    x = _need(x, "embeddings", torch.float32)
    n, d = x.shape
    cid = _need(cluster_ids, "cluster_ids", torch.int64, (n,))
    qs = _need(queries, "queries", torch.int64)
    out = torch.empty(5, dtype=torch.float32, device=x.device)
    ws = _workspace(lib.hept_knn_metrics_workspace_bytes(n, qs.numel()), x)
    _lib.check(lib.hept_knn_metrics(_ptr(x), n, d, _ptr(cid), _ptr(qs), qs.numel(), 1 if cosine else 0, K, _ptr(out), _ptr(ws),
                                    ws.numel(), _stream(x)), "hept_knn_metrics")
    return out


# ------------------------------------------------------------------------------- a13..a17 preparation
@_on_device
def prepare_batched(coords, batch, offsets, num_events: int, n_raw: int, n_pad: int, max_event: int, regions_h, block_size: int,
                    want_int32: bool = True, code_bits: int = 0):
    """example/ flavour of prepare_input on the library's kernels.  ``offsets`` = int32 device tensor
    [event_start (E+1) | pad_start (E+1)].  -> combined_shifts (TH, n_pad) int64, the same as int32 (or None),
    take (n_pad) int64, is_real (n_pad) bool, coords_pad (n_pad, C)."""
    lib = _lib.load()
    coords = _need(coords, "coords", torch.float32, (n_raw, coords.shape[-1]))
    batch = _need(batch, "batch", torch.int64, (n_raw,))
    offsets = _need(offsets, "offsets", torch.int32, (2 * (num_events + 1),))
    th = regions_h.shape[1]
    regions_h = _need(regions_h, "regions_h", torch.float32, (2, th))
    dev, c = coords.device, coords.shape[1]
    shifts = torch.empty(th, n_pad, dtype=torch.int64, device=dev)
    shifts32 = torch.empty(th, n_pad, dtype=torch.int32, device=dev) if want_int32 else None
    take = torch.empty(n_pad, dtype=torch.int64, device=dev)
    real = torch.empty(n_pad, dtype=torch.uint8, device=dev)
    coords_pad = torch.empty(n_pad, c, dtype=torch.float32, device=dev)
    ws = _workspace(lib.hept_prepare_batched_workspace_bytes(n_raw, num_events, max_event), coords)
    ev_start = C.c_void_p(offsets.data_ptr())
    pad_start = C.c_void_p(offsets.data_ptr() + 4 * (num_events + 1))
    _lib.check(lib.hept_prepare_batched(_ptr(coords), c, _ptr(batch), ev_start, pad_start, num_events, n_raw, n_pad, max_event,
                                        _ptr(regions_h), th, block_size, int(code_bits), _ptr(shifts), _ptr(shifts32), _ptr(take), _ptr(real),
                                        _ptr(coords_pad), _ptr(ws), ws.numel(), _stream(coords)), "hept_prepare_batched")
    return shifts, shifts32, take, real.view(torch.bool), coords_pad


@_on_device
def prepare_single(coords, n_pad: int, regions_h):
    """src/ flavour of prepare_input.  -> coords_pad (n_pad, C), region_eta, region_phi (TH, n_pad) float32."""
    lib = _lib.load()
    coords = _need(coords, "coords", torch.float32)
    n_raw, c = coords.shape
    th = regions_h.shape[1]
    regions_h = _need(regions_h, "regions_h", torch.float32, (2, th))
    dev = coords.device
    coords_pad = torch.empty(n_pad, c, dtype=torch.float32, device=dev)
    eta = torch.empty(th, n_pad, dtype=torch.float32, device=dev)
    phi = torch.empty(th, n_pad, dtype=torch.float32, device=dev)
    ws = _workspace(lib.hept_prepare_single_workspace_bytes(n_pad), coords)
    _lib.check(lib.hept_prepare_single(_ptr(coords), c, n_raw, n_pad, _ptr(regions_h), th, _ptr(coords_pad), _ptr(eta),
                                       _ptr(phi), _ptr(ws), ws.numel(), _stream(coords)), "hept_prepare_single")
    return coords_pad, eta, phi


# --------------------------------------------------------------------------------------- whole path
@_on_device
def attention_fwd(d: Dims, q, k, v, coords, w_rpe_weight, K: int, alpha, combined_shifts=None, region_indices=None,
                  regions_h=None):
    """a3..a12 in one native call -> (out_pre (N,H*D), den_sum (N,H), scale (H,C), positions (2,T,H,N) int32)."""
    lib = _lib.load()
    q = _need(q, "query", torch.float32, (d.N, d.H * d.D))
    k = _need(k, "key", torch.float32, (d.N, d.H * d.D))
    v = _need(v, "value", torch.float32, (d.N, d.H * d.D))
    coords = _need(coords, "coords", torch.float32, (d.N, d.C))
    w = _need(w_rpe_weight, "w_rpe.weight", torch.float32, (d.H * d.D, (d.C - 1) * K))
    alpha = _need(alpha, "e2lsh.alpha", torch.float32, (d.H, d.E, d.T))
    sh = eta = phi = rh = None
    sh32 = combined_shifts is not None and combined_shifts.dtype == torch.int32
    if combined_shifts is not None:
        sh = _need(combined_shifts, "combined_shifts", torch.int32 if sh32 else torch.int64, (d.T, d.H, d.N))
    else:
        eta = _need(region_indices[0], "region_indices[0]", torch.float32, (d.T * d.H, d.N))
        phi = _need(region_indices[1], "region_indices[1]", torch.float32, (d.T * d.H, d.N))
        rh = _need(regions_h, "regions_h", torch.float32, (2, d.T * d.H))
    dev = q.device
    scale = torch.empty(d.H, d.C, dtype=torch.float32, device=dev)
    pos = torch.empty(2, d.T, d.H, d.N, dtype=torch.int32, device=dev)
    out_pre = torch.empty(d.N, d.H * d.D, dtype=torch.float32, device=dev)
    den = torch.empty(d.N, d.H, dtype=torch.float32, device=dev)
    s = d.struct()
    ws = _workspace(lib.hept_attention_fwd_workspace_bytes(C.byref(s)), q)
    if sh32:
        _lib.check(lib.hept_attention_fwd_shifts32(C.byref(s), _ptr(q), _ptr(k), _ptr(v), _ptr(coords), _ptr(w), K, _ptr(alpha),
                                                   _ptr(sh), _ptr(scale), _ptr(pos), _ptr(out_pre), _ptr(den), _ptr(ws),
                                                   ws.numel(), _stream(q)), "hept_attention_fwd_shifts32")
        return out_pre, den, scale, pos
    _lib.check(lib.hept_attention_fwd(C.byref(s), _ptr(q), _ptr(k), _ptr(v), _ptr(coords), _ptr(w), K, _ptr(alpha),
                                      _ptr(sh), _ptr(eta), _ptr(phi), _ptr(rh), _ptr(scale), _ptr(pos), _ptr(out_pre),
                                      _ptr(den), _ptr(ws), ws.numel(), _stream(q)), "hept_attention_fwd")
    return out_pre, den, scale, pos


@_on_device
def attention_bwd(d: Dims, q, k, v, coords, scale, positions, out_pre, den_sum, d_out_pre):
    """-> dq, dk, dv (N, H*D), dscale (H, C)."""
    lib = _lib.load()
    q = _need(q, "query", torch.float32, (d.N, d.H * d.D))
    k = _need(k, "key", torch.float32, (d.N, d.H * d.D))
    v = _need(v, "value", torch.float32, (d.N, d.H * d.D))
    coords = _need(coords, "coords", torch.float32, (d.N, d.C))
    scale = _need(scale, "scale", torch.float32, (d.H, d.C))
    pos = _need(positions, "positions", torch.int32, (2, d.T, d.H, d.N))
    out_pre = _need(out_pre, "out_pre", torch.float32, (d.N, d.H * d.D))
    den_sum = _need(den_sum, "den_sum", torch.float32, (d.N, d.H))
    g = _need(d_out_pre, "d_out_pre", torch.float32, (d.N, d.H * d.D))
    dq, dk, dv = torch.empty_like(q), torch.empty_like(q), torch.empty_like(q)
    dscale = torch.empty(d.H, d.C, dtype=torch.float32, device=q.device)
    s = d.struct()
    ws = _workspace(lib.hept_attention_bwd_workspace_bytes(C.byref(s)), q)
    _lib.check(lib.hept_block_attention_bwd(C.byref(s), _ptr(q), _ptr(k), _ptr(v), _ptr(coords), _ptr(scale), _ptr(pos),
                                            _ptr(out_pre), _ptr(den_sum), _ptr(g), _ptr(dq), _ptr(dk), _ptr(dv),
                                            _ptr(dscale), _ptr(ws), ws.numel(), _stream(q)), "hept_block_attention_bwd")
    return dq, dk, dv, dscale

"""ctypes binding of libhept_sm100.so (include/hept_b200.h).

There is no fallback: if the shared library is missing or a call fails, a RuntimeError is raised.
The library is built in-tree by ``python -m hept_b200.build`` (or ``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HEPT_LIB") or os.path.join(_HERE, "libhept_sm100.so")   # HEPT_LIB: tools/pipeline_trace.py
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "hept_b200.h")

HEPT_OK, HEPT_EINVAL, HEPT_EUNSUPPORTED, HEPT_ECUDA, HEPT_EWORKSPACE = 0, -1, -2, -3, -4


class Shape(C.Structure):
    """struct hept_shape"""

    _fields_ = [(n, C.c_int32) for n in ("N", "H", "D", "C", "T", "B", "raw_size")]


_p, _i32, _sz = C.c_void_p, C.c_int32, C.c_size_t
_SP = C.POINTER(Shape)

# name -> (restype, argtypes); mirrors include/hept_b200.h one to one
SIGNATURES = {
    "hept_abi_version": (C.c_int, []),
    "hept_last_error": (C.c_char_p, []),
    "hept_shape_supported": (C.c_int, [_i32, _i32, _i32]),
    "hept_coord_scale_fwd": (C.c_int, [_p, _i32, _i32, _i32, _i32, _p, _p]),
    "hept_coord_scale_bwd": (C.c_int, [_p, _p, _p, _i32, _i32, _i32, _i32, _p, _p]),
    "hept_hash_workspace_bytes": (_sz, [_SP]),
    "hept_hash_project": (C.c_int, [_SP, _p, _p, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "hept_keys_from_packed_shifts": (C.c_int, [_SP, _p, _p, _p, _p, _p]),
    "hept_keys_from_region_indices": (C.c_int, [_SP, _p, _p, _p, _p, _p, _p, _p]),
    "hept_argsort_workspace_bytes": (_sz, [_i32, _i32]),
    "hept_segmented_argsort": (C.c_int, [_p, _i32, _i32, _p, _p, _sz, _p]),
    "hept_hat_coords": (C.c_int, [_SP, _p, _p, _p, _p]),
    "hept_block_attention_fwd": (C.c_int, [_SP, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "hept_or_combine": (C.c_int, [_SP, _p, _p, _p, _p]),
    "hept_out_linear_fwd": (C.c_int, [_SP, _p, _p, _p, _p, _p]),
    "hept_out_linear_bwd_workspace_bytes": (_sz, [_SP]),
    "hept_out_linear_bwd": (C.c_int, [_SP, _p, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "hept_attention_bwd_workspace_bytes": (_sz, [_SP]),
    "hept_block_attention_bwd": (C.c_int, [_SP] + [_p] * 14 + [_sz, _p]),
    "hept_attention_fwd_workspace_bytes": (_sz, [_SP]),
    "hept_attention_fwd": (C.c_int, [_SP, _p, _p, _p, _p, _p, _i32, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "hept_prepare_batched_workspace_bytes": (_sz, [_i32, _i32, _i32]),
    "hept_prepare_batched": (C.c_int, [_p, _i32, _p, _p, _p, _i32, _i32, _i32, _i32, _p, _i32, _i32, _i32, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "hept_prepare_single_workspace_bytes": (_sz, [_i32]),
    "hept_prepare_single": (C.c_int, [_p, _i32, _i32, _i32, _p, _i32, _p, _p, _p, _p, _sz, _p]),
    "hept_keys_from_packed_shifts32": (C.c_int, [_SP, _p, _p, _p, _p, _p]),
    "hept_attention_fwd_shifts32": (C.c_int, [_SP, _p, _p, _p, _p, _p, _i32, _p, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "hept_attn_qkv_supported": (C.c_int, [_i32, _i32]),
    "hept_attn_qkv_fwd": (C.c_int, [_p, _p, _p, _p, _p, _p, _i32, _i32, _i32, C.c_float, _p, _p, _p, _p, _p, _p]),
    "hept_attn_qkv_bwd_workspace_bytes": (_sz, [_i32, _i32, _i32]),
    "hept_attn_qkv_bwd": (C.c_int, [_p, _p, _p, _p, _p, _p, _p, _i32, _i32, _i32, C.c_float, _p, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "hept_layer_norm_supported": (C.c_int, [_i32]),
    "hept_layer_norm_fwd": (C.c_int, [_p, _p, _p, _i32, _i32, C.c_float, _p, _p, _p]),
    "hept_layer_norm_bwd_workspace_bytes": (_sz, [_i32, _i32]),
    "hept_layer_norm_bwd": (C.c_int, [_p, _p, _p, _p, _i32, _i32, _p, _p, _p, _p, _sz, _p]),
    "hept_infonce_saved_bytes": (_sz, [_i32, C.c_int64]),
    "hept_infonce_workspace_bytes": (_sz, [_i32, C.c_int64, _i32]),
    "hept_infonce_fwd": (C.c_int, [_p, _i32, _i32, _p, C.c_int64, _p, _p, _p, C.c_float, _i32, C.c_float, _p, _p, _sz, _p, _sz, _p]),
    "hept_infonce_bwd": (C.c_int, [_p, _i32, _i32, _p, C.c_int64, _i32, C.c_float, _p, _p, _sz, _p, _p, _sz, _p]),
    "hept_knn_metrics_workspace_bytes": (_sz, [_i32, _i32]),
    "hept_knn_metrics": (C.c_int, [_p, _i32, _i32, _p, _p, _i32, _i32, _i32, _p, _p, _sz, _p]),
    "hept_p2p_flag_bytes": (_sz, []),
    "hept_p2p_allreduce": (C.c_int, [_p, _i32, _i32, C.c_int64, C.c_uint32, C.c_float, _p, _p, _p]),
    "hept_launch_count": (C.c_int, [C.c_int]),
    "hept_set_bwd_stage_mask": (None, [C.c_int]),
    "hept_set_engine": (None, [C.c_int]),
    "hept_get_engine": (C.c_int, []),
    "hept_set_bwd_variant": (None, [C.c_int]),
    "hept_set_sort_variant": (None, [C.c_int]),
    "hept_get_sort_variant": (C.c_int, []),
    "hept_get_bwd_variant": (C.c_int, []),
}

_lib = None
_lock = threading.Lock()


class HeptLibraryError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load the native library once; raise loudly if it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise HeptLibraryError(
                    f"{LIB_PATH} is missing: hept_b200 has no CPU or PyTorch fallback. "
                    "Build it with `python -m hept_b200.build` (needs nvcc, targets sm_100a)."
                )
            lib = C.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)  # AttributeError here == header and library disagree
                fn.restype, fn.argtypes = res, args
            _lib = lib
    return _lib


def check(rc: int, what: str) -> None:
    if rc != HEPT_OK:
        msg = load().hept_last_error().decode("utf-8", "replace")
        kind = {HEPT_EINVAL: ValueError, HEPT_EUNSUPPORTED: NotImplementedError}.get(rc, RuntimeError)
        raise kind(f"{what} failed (code {rc}): {msg}")

"""ctypes binding of the test-only library tests/native/libhept_umma_test.so (tcgen05 building-block self-tests)."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhept_umma_test.so")
_p = C.c_void_p
_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `make -C tests/native` (or __graft_entry__.build())")
        lib = C.CDLL(LIB_PATH)
        lib.hept_debug_last_error.restype = C.c_char_p
        lib.hept_debug_umma_selftest.argtypes = [_p, _p, _p, _p, _p, C.c_int, _p]
        lib.hept_debug_umma_symmetry.argtypes = [_p, _p, _p, _p, _p]
        lib.hept_debug_umma_timing.argtypes = [C.c_int, C.c_int, _p, _p]
        _lib = lib
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"{what} failed (code {rc}): {load().hept_debug_last_error().decode()}")

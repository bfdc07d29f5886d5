"""tcgen05 / TMEM building blocks on exactly representable data (small integers survive the tf32 operand
truncation and fp32 accumulation exactly), so a wrong descriptor bit or swizzle shows up as inequality."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_umma_selftest_exact_on_small_integers():
    from hept_b200 import _lib

    lib = _lib.load()
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    A = torch.randint(-4, 5, (128, 32), generator=g).float()
    B = torch.randint(-4, 5, (112, 32), generator=g).float()
    V = torch.randint(-3, 4, (112, 32), generator=g).float()
    S = torch.empty(128, 112, device=dev)
    O = torch.empty(128, 32, device=dev)
    Ad, Bd, Vd = A.to(dev), B.to(dev), V.to(dev)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    rc = lib.hept_debug_umma_selftest(p(Ad), p(Bd), p(Vd), p(S), p(O), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(rc, "hept_debug_umma_selftest")
    torch.cuda.synchronize()
    S_ref = A.double() @ B.double().T
    O_ref = S_ref @ V.double()
    assert torch.equal(S.cpu().double(), S_ref), (S.cpu()[:2, :8], S_ref[:2, :8])
    assert torch.equal(O.cpu().double(), O_ref), (O.cpu()[:2, :8], O_ref[:2, :8])

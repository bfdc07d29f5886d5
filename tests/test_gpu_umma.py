"""tcgen05 / TMEM building blocks on exactly representable data (small integers survive the tf32 operand
truncation and fp32 accumulation exactly), so a wrong descriptor bit or swizzle shows up as inequality."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kmajor_base32", [0])   # 1 (K-major operands in the MN-major operand's SWIZZLE_128B_BASE32B
def test_umma_selftest_exact_on_small_integers(kmajor_base32):   # image) faults on sm_100a: the layouts cannot be shared
    from tests import native as _lib

    lib = _lib.load()
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    A = torch.randint(-4, 5, (128, 32), generator=g).float()
    B = torch.randint(-4, 5, (112, 32), generator=g).float()
    V = torch.randint(-3, 4, (112, 32), generator=g).float()
    S = torch.empty(128, 112, device=dev)
    O = torch.empty(128, 96, device=dev)
    Ad, Bd, Vd = A.to(dev), B.to(dev), V.to(dev)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    rc = lib.hept_debug_umma_selftest(p(Ad), p(Bd), p(Vd), p(S), p(O), kmajor_base32,
                                      ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(rc, "hept_debug_umma_selftest")
    torch.cuda.synchronize()
    S_ref = A.double() @ B.double().T
    O_ref = S_ref @ V.double()
    assert torch.equal(S.cpu().double(), S_ref), (S.cpu()[:2, :8], S_ref[:2, :8])
    assert torch.equal(O.cpu()[:, :32].double(), O_ref), (O.cpu()[:2, :8], O_ref[:2, :8])
    # one N = 64 MMA per k-step over two 32-column tiles LBO apart: [S V | S (2 V + 1)]
    O2_ref = torch.cat([O_ref, S_ref @ (2 * V.double() + 1)], dim=1)
    assert torch.equal(O.cpu()[:, 32:].double(), O2_ref), (O.cpu()[:2, 32:40], O2_ref[:2, :8])


def test_tf32_mma_is_symmetric_under_operand_exchange():
    """x_i . y_j from MMA(X, Y) and from MMA(Y, X) on random fp32 data: bit-identical?  (Decides whether a tensor-core
    backward can recompute S and S^T in two kernels and still see the same dS; recorded, not required.)"""
    import json
    import os

    from tests import native as _lib

    lib = _lib.load()
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(1)
    X = (torch.randn(112, 32, generator=g) * 3).to(dev)
    Y = (torch.randn(112, 32, generator=g) * 3).to(dev)
    A = torch.empty(128, 112, device=dev)
    B = torch.empty(128, 112, device=dev)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    _lib.check(lib.hept_debug_umma_symmetry(p(X), p(Y), p(A), p(B), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)),
               "hept_debug_umma_symmetry")
    torch.cuda.synchronize()
    s_xy, s_yx = A[:112].cpu(), B[:112].cpu().T
    ref = X.cpu().double() @ Y.cpu().double().T
    same = bool(torch.equal(s_xy, s_yx))
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/umma_symmetry.json", "w") as f:
        json.dump({"bitwise_symmetric": same, "max_abs_diff": float((s_xy - s_yx).abs().max()),
                   "tf32_rel_err": float((s_xy.double() - ref).norm() / ref.norm())}, f)
    assert float((s_xy.double() - ref).norm() / ref.norm()) < 2e-3      # plain tf32 product: ~2^-11 relative

"""The caller's LayerNorms (Attn block norm2, example/transformer.py:163; the model head's 256-wide norms) on the library's
kernels (hept_layer_norm_fwd / _bwd through the C ABI) against torch.nn.functional.layer_norm evaluated in float64, with the
float32 evaluation of the same torch op as the yardstick:
    err(ours vs fp64) <= 2.5 * err(torch fp32 vs fp64) + floor,   floor = 2e-6 (output, dx) / 1e-5 (parameter gradients)."""
import pytest
import torch
import torch.nn.functional as F

from tests.helpers import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _reference(x, g, b, dy, dtype, eps=1e-5):
    t = lambda a: a.detach().to(dtype).clone().requires_grad_(True)
    x, g, b = t(x), t(g), t(b)
    y = F.layer_norm(x, (x.shape[-1],), g, b, eps)
    (y * dy.to(dtype)).sum().backward()
    return {"y": y.detach(), "dx": x.grad, "dw": g.grad, "db": b.grad}


@pytest.mark.parametrize("d", [24, 256, 12, 64, 100])
@pytest.mark.parametrize("n", [1, 7, 33, 1300, 60000])
def test_layer_norm_forward_backward(n, d):
    from hept_b200 import ops

    gen = torch.Generator().manual_seed(1000 * d + n)
    x = torch.randn(n, d, generator=gen) * 1.3 + 0.4
    g, b = 1 + 0.2 * torch.randn(d, generator=gen), 0.1 * torch.randn(d, generator=gen)
    dy = torch.randn(n, d, generator=gen)
    r32, r64 = _reference(x, g, b, dy, torch.float32), _reference(x, g, b, dy, torch.float64)
    dev = lambda a: a.to(DEV)
    y, mr = ops.layer_norm_fwd(dev(x), dev(g), dev(b), 1e-5)
    dx, dw, db = ops.layer_norm_bwd(dev(x), mr, dev(g), dev(dy))
    bad = []
    for key, val in {"y": y, "dx": dx, "dw": dw, "db": db}.items():
        floor = 2e-6 if key in ("y", "dx") else 1e-5
        e_o, e_r = rel_err(val.cpu(), r64[key]), rel_err(r32[key], r64[key])
        if not e_o <= 2.5 * e_r + floor:
            bad.append((key, e_o, e_r))
    assert not bad, bad
    again = ops.layer_norm_bwd(dev(x), mr, dev(g), dev(dy))          # fixed-order reductions: bit for bit
    for a, c in zip((dx, dw, db), again):
        assert torch.equal(a, c)


def test_layer_norm_module_is_a_drop_in_for_nn_layer_norm():
    """Same state_dict keys and parameters as nn.LayerNorm; 3-D inputs; the gradient reaches x, weight and bias."""
    from hept_b200.layers import LayerNorm

    torch.manual_seed(3)
    ref = torch.nn.LayerNorm(24).to(DEV)
    with torch.no_grad():
        ref.weight.uniform_(0.5, 1.5); ref.bias.uniform_(-0.2, 0.2)
    mine = LayerNorm(24).to(DEV)
    mine.load_state_dict(ref.state_dict(), strict=True)
    x = torch.randn(5, 130, 24, device=DEV)
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    ya, yb = mine(xa), ref(xb)
    assert torch.allclose(ya, yb, rtol=1e-5, atol=1e-6)
    w = torch.randn_like(ya)
    (ya * w).sum().backward(); (yb * w).sum().backward()
    assert torch.allclose(xa.grad, xb.grad, rtol=1e-4, atol=1e-5)
    assert torch.allclose(mine.weight.grad, ref.weight.grad, rtol=1e-4, atol=1e-4)
    assert torch.allclose(mine.bias.grad, ref.bias.grad, rtol=1e-4, atol=1e-4)


def test_unsupported_width_is_refused_by_the_abi():
    from hept_b200 import ops

    assert not ops.layer_norm_supported(258) and not ops.layer_norm_supported(10) and ops.layer_norm_supported(256)
    with pytest.raises(Exception):
        ops.layer_norm_fwd(torch.zeros(4, 10, device=DEV), torch.ones(10, device=DEV), torch.zeros(10, device=DEV), 1e-5)


@pytest.mark.parametrize("shape", [(60000, 256, 256), (60000, 24, 24), (5000, 120, 12), (4097, 15, 24)])
def test_tall_linear_matches_nn_linear(shape):
    """The split-K weight gradient of hept_b200.layers.Linear against nn.Linear's own autograd (fp64 yardstick, 2.5x rule)."""
    from hept_b200.layers import Linear

    n, din, dout = shape
    torch.manual_seed(n + din)
    ref = torch.nn.Linear(din, dout).to(DEV)
    mine = Linear(din, dout).to(DEV)
    mine.load_state_dict(ref.state_dict(), strict=True)
    ref64 = torch.nn.Linear(din, dout).to(DEV).double()
    ref64.load_state_dict({k: v.double() for k, v in ref.state_dict().items()})
    x = torch.randn(n, din, device=DEV)
    g = torch.randn(n, dout, device=DEV)
    outs = {}
    for name, mod, xx, gg in (("mine", mine, x, g), ("ref", ref, x, g), ("ref64", ref64, x.double(), g.double())):
        xi = xx.clone().requires_grad_(True)
        y = mod(xi)
        (y * gg).sum().backward()
        outs[name] = {"y": y.detach(), "dx": xi.grad, "dw": mod.weight.grad, "db": mod.bias.grad}
    for key in ("y", "dx", "dw", "db"):
        e_o = rel_err(outs["mine"][key].cpu(), outs["ref64"][key].cpu())
        e_r = rel_err(outs["ref"][key].cpu(), outs["ref64"][key].cpu())
        assert e_o <= 2.5 * e_r + 2e-6, (key, e_o, e_r)

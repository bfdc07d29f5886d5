"""CUDA-graph replay of the module call (hept_b200/graphed.py): same bits as the eager call."""
import pytest
import torch

from hept_b200 import synthetic

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _problem(n_raw, seed):
    from hept_b200 import HEPTAttention, prepare

    cfg = dict(synthetic.TRACKING)
    coords, batch = synthetic.batched_cloud([n_raw], 6, seed)
    params = synthetic.module_params(cfg, seed)
    helper = {"block_size": 100, "regions": params["regions"].to(DEV), "num_heads": 8}
    _, kw, _ = prepare.prepare_input(torch.zeros(n_raw, 1, device=DEV), coords.to(DEV), batch.to(DEV), helper, sizes=[n_raw])
    n = kw["coords"].shape[0]
    mod = HEPTAttention(30, **cfg)
    mod.load_state_dict({k: params[k] for k in ("out_linear.weight", "out_linear.bias", "e2lsh.alpha")}, strict=True)
    w_rpe = torch.nn.Linear(50, 192)
    w_rpe.load_state_dict({"weight": params["w_rpe.weight"], "bias": params["w_rpe.bias"]})
    return cfg, mod.to(DEV), w_rpe.to(DEV), kw, n


@pytest.mark.parametrize("n_raw", [6037, 60000])
def test_graphed_attention_replays_the_eager_call_bit_for_bit(n_raw):
    from hept_b200.graphed import graphed_attention

    cfg, mod, w_rpe, kw, n = _problem(n_raw, 3)
    qkv = [t.to(DEV) for t in synthetic.qkv(n, cfg, 3)]
    g = torch.randn(n, 24, generator=torch.Generator().manual_seed(1)).to(DEV)
    q, k, v = (t.clone().requires_grad_(True) for t in qkv)
    out = mod(q, k, v, w_rpe=w_rpe, coords=kw["coords"], combined_shifts=kw["combined_shifts32"])
    out.backward(g)
    want = [out.detach().clone(), q.grad.clone(), k.grad.clone(), v.grad.clone(), w_rpe.weight.grad.clone(),
            mod.out_linear.weight.grad.clone()]
    for p in (w_rpe.weight, mod.out_linear.weight, mod.out_linear.bias):
        p.grad = None
    step = graphed_attention(mod, w_rpe, q.detach().requires_grad_(True), k.detach().requires_grad_(True),
                             v.detach().requires_grad_(True), kw["coords"], kw["combined_shifts32"])
    for rep in range(2):       # the second replay runs on new data placed in the same buffers
        src = qkv if rep == 0 else [t.to(DEV) for t in synthetic.qkv(n, cfg, 4)]
        q2, k2, v2 = (t.clone().requires_grad_(True) for t in src)
        for p in (w_rpe.weight, mod.out_linear.weight, mod.out_linear.bias):
            p.grad = None
        out2 = step(q2, k2, v2, kw["coords"], kw["combined_shifts32"])
        out2.backward(g)
        if rep == 0:
            got = [out2.detach(), q2.grad, k2.grad, v2.grad, w_rpe.weight.grad, mod.out_linear.weight.grad]
            for a, b in zip(got, want):
                assert torch.equal(a, b)
        else:
            q3, k3, v3 = (t.clone().requires_grad_(True) for t in src)
            for p in (w_rpe.weight, mod.out_linear.weight, mod.out_linear.bias):
                p.grad = None
            out3 = mod(q3, k3, v3, w_rpe=w_rpe, coords=kw["coords"], combined_shifts=kw["combined_shifts32"])
            assert torch.equal(out2.detach(), out3.detach())


def test_graphed_pileup_inference_matches_eager():
    from hept_b200.graphed import GraphedInference
    from hept_b200.model import Transformer

    cfg = {k: v for k, v in synthetic.PILEUP.items() if k != "coords_dim"}
    torch.manual_seed(6)
    m = Transformer(in_dim=8, coords_dim=4, task="pileup", flavour="src", **cfg).eval().to(DEV)
    n = 10000
    coords = synthetic.point_cloud(n, 4, 8).to(DEV)
    x = torch.cat([torch.randn(n, 7) * 0.5, torch.randint(0, 7, (n, 1)).float()], dim=1).to(DEV)
    with torch.no_grad():
        want = m(x, coords).clone()
    gi = GraphedInference(m, (x, coords))
    assert torch.equal(gi.run(x, coords), want)
    coords2 = synthetic.point_cloud(n, 4, 9).to(DEV)
    with torch.no_grad():
        want2 = m(x, coords2).clone()
    assert torch.equal(gi.run(x, coords2), want2)

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
    params = (w_rpe.weight, mod.out_linear.weight, mod.out_linear.bias)
    # capture first: an autograd graph of an earlier eager call that is still alive would tie the capture to the legacy stream
    step = graphed_attention(mod, w_rpe, *(t.clone().requires_grad_(True) for t in qkv), kw["coords"], kw["combined_shifts32"])

    def run(fn, src):
        q, k, v = (t.clone().requires_grad_(True) for t in src)
        for p in params:
            p.grad = None
        out = fn(q, k, v)
        out.backward(g)
        res = [out.detach().clone(), q.grad.clone(), k.grad.clone(), v.grad.clone()] + [p.grad.clone() for p in params]
        del out
        return res

    eager = lambda q, k, v: mod(q, k, v, w_rpe=w_rpe, coords=kw["coords"], combined_shifts=kw["combined_shifts32"])
    replay = lambda q, k, v: step(q, k, v, kw["coords"], kw["combined_shifts32"])
    for seed in (3, 4):        # the second replay runs on new data placed in the same captured buffers
        src = [t.to(DEV) for t in synthetic.qkv(n, cfg, seed)]
        got, want = run(replay, src), run(eager, src)
        for a, b in zip(got, want):
            assert torch.equal(a, b)


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

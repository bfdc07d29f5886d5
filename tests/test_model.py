"""The Transformer wrapper around the hot path (hept_b200/model.py): checkpoint compatibility on CPU, and on the
GPU the whole model (4 attention layers, forward + backward) against the same wrapper driven by the CPU oracle."""
import os

import pytest
import torch

from hept_b200 import synthetic
from hept_b200.model import NodeMLP, Transformer

CKPT = "/root/reference/example/ckpt/tracking-60k-model.pt"
TRACKING = {k: v for k, v in synthetic.TRACKING.items() if k != "coords_dim"}      # the reference's model_kwargs
PILEUP = {k: v for k, v in synthetic.PILEUP.items() if k != "coords_dim"}


def test_state_dict_layout_matches_reference_model():
    m = Transformer(in_dim=15, coords_dim=6, num_classes=0, **TRACKING)
    keys = set(m.state_dict())
    assert {"regions", "W.weight", "feat_encoder.0.weight", "feat_encoder.2.bias", "mlp_out.lins.4.weight",
            "mlp_out.norms.3.bias", "attns.3.attn.e2lsh.alpha", "attns.0.attn.out_linear.weight", "attns.2.w_rpe.weight",
            "attns.1.w_rpe.bias", "attns.0.norm1.weight", "attns.0.ff.2.weight", "attns.0.w_q.weight"} <= keys
    assert len(keys) == 88                                   # 88 tensors in the shipped checkpoint
    assert sum(p.numel() for p in m.state_dict().values()) == 329_364   # 329,364 parameters
    assert tuple(m.regions.shape) == (3, 2, 8) and not m.regions.requires_grad
    p = Transformer(in_dim=8, coords_dim=4, task="pileup", flavour="src", **PILEUP)
    assert tuple(p.feat_encoder[0].weight.shape) == (24, 17) and tuple(p.attns[0].w_rpe.weight.shape) == (192, 30)


@pytest.mark.skipif(not os.path.exists(CKPT), reason="reference checkpoint only exists in the build container")
def test_reference_checkpoint_loads_strictly():
    sd = torch.load(CKPT, map_location="cpu")
    m = Transformer(in_dim=15, coords_dim=6, num_classes=0, **TRACKING)
    m.load_state_dict(sd, strict=True)
    assert torch.equal(m.attns[2].attn.e2lsh.alpha, sd["attns.2.attn.e2lsh.alpha"])


def test_node_mlp_is_lin_norm_tanh():
    torch.manual_seed(0)
    mlp = NodeMLP(12, 256, 12, 5)
    x = torch.randn(7, 12)
    y = x
    for i in range(4):
        y = torch.tanh(torch.nn.functional.layer_norm(mlp.lins[i](y), (256,), mlp.norms[i].weight, mlp.norms[i].bias))
    assert torch.allclose(mlp(x), mlp.lins[4](y))


class OracleAttention(torch.nn.Module):
    """Same parameters and call surface as HEPTAttention, computed by the CPU oracle (tests only)."""

    def __init__(self, hash_dim, **kw):
        super().__init__()
        from hept_b200.attention import E2LSH

        self.cfg = kw
        self.out_linear = torch.nn.Linear(kw["num_heads"] * kw["h_dim"], kw["h_dim"])
        self.e2lsh = E2LSH(kw["n_hashes"], kw["num_heads"], hash_dim)

    def forward(self, q, k, v, **kwargs):
        from oracle import hept_oracle as O

        flavour = ({"combined_shifts": kwargs["combined_shifts"]} if "combined_shifts" in kwargs else
                   {"raw_size": kwargs["raw_size"], "regions_h": kwargs["regions_h"], "region_indices": kwargs["region_indices"]})
        return O.attention_forward(q, k, v, out_weight=self.out_linear.weight, out_bias=self.out_linear.bias,
                                   w_rpe_weight=kwargs["w_rpe"].weight, alpha=self.e2lsh.alpha, coords=kwargs["coords"],
                                   block_size=self.cfg["block_size"], num_heads=self.cfg["num_heads"],
                                   dim_per_head=self.cfg["h_dim"], num_w_per_dist=self.cfg["num_w_per_dist"], **flavour)


class OraclePrepare:
    """prepare_input / prepare_input_single of the oracle behind the product module's call surface (tests only)."""

    @staticmethod
    def prepare_input(x, coords, batch, helper):
        from oracle import hept_oracle as O

        return O.prepare_batched(x, coords, batch, helper["regions"], helper["block_size"], helper["num_heads"], stable=True)

    @staticmethod
    def prepare_input_single(x, coords, helper):
        from oracle import hept_oracle as O

        return O.prepare_single_event(x, coords, helper["regions"], helper["block_size"])


@pytest.mark.gpu
def test_whole_model_forward_backward_against_oracle_backed_model():
    cfg = dict(TRACKING)
    torch.manual_seed(5)
    ours = Transformer(in_dim=15, coords_dim=6, **cfg).eval()
    ref = Transformer(in_dim=15, coords_dim=6, attn_cls=OracleAttention, prepare_impl=OraclePrepare, **cfg).eval()
    ref.load_state_dict(ours.state_dict(), strict=True)
    sizes = [830, 411, 57]
    coords, batch = synthetic.batched_cloud(sizes, 6, 3)
    x = torch.randn(coords.shape[0], 15, generator=torch.Generator().manual_seed(1)) * 0.5
    dev = torch.device("cuda:0")
    ours = ours.to(dev)
    out = ours(x.to(dev), coords.to(dev), batch.to(dev))
    want = ref(x, coords, batch)
    assert out.shape == want.shape == (sum(sizes), 12)
    row = (out.cpu() - want).norm(dim=1) / want.norm(dim=1).clamp_min(1e-12)
    assert float(row.median()) < 2e-4 and float((row > 1e-2).float().mean()) < 0.05
    g = torch.randn(want.shape, generator=torch.Generator().manual_seed(2))
    out.backward(g.to(dev))
    want.backward(g)
    for name in ("W.weight", "attns.0.w_q.weight", "attns.3.w_rpe.weight", "feat_encoder.0.weight", "attns.1.attn.out_linear.bias"):
        a = dict(ours.named_parameters())[name].grad.cpu()
        b = dict(ref.named_parameters())[name].grad
        assert float((a - b).norm() / b.norm()) < 5e-3, name
    assert dict(ours.named_parameters())["attns.0.w_rpe.bias"].grad is None


@pytest.mark.gpu
def test_pileup_model_inference_src_flavour():
    cfg = dict(PILEUP)
    torch.manual_seed(6)
    m = Transformer(in_dim=8, coords_dim=4, task="pileup", flavour="src", **cfg).eval().to("cuda:0")
    n = 4321
    coords = synthetic.point_cloud(n, 4, 8)
    x = torch.cat([torch.randn(n, 7) * 0.5, torch.randint(0, 7, (n, 1)).float()], dim=1)
    ref = Transformer(in_dim=8, coords_dim=4, task="pileup", flavour="src", attn_cls=OracleAttention, prepare_impl=OraclePrepare,
                      **cfg).eval()
    ref.load_state_dict(m.state_dict(), strict=True)
    with torch.no_grad():
        y = m(x.cuda(), coords.cuda())
        want = ref(x, coords)
    assert y.shape == want.shape == (n, 1) and bool(((y > 0) & (y < 1)).all())
    # BASELINE.json configs[2]: forward-only inference of the pileup model against the oracle-backed twin (same weights):
    # four attention layers deep, a handful of rows sit in blocks whose tie order differs
    err = (y.cpu() - want).abs().squeeze(1)
    assert float(err.median()) < 2e-5 and float((err > 1e-3).float().mean()) < 0.05

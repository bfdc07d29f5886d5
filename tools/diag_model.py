import sys, os
sys.path.insert(0, "/root/repo")
import torch
from hept_b200 import synthetic, _lib, prepare
from hept_b200.model import Transformer
from tests.test_model import OracleAttention, OraclePrepare, TRACKING
lib = _lib.load()
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
helper = dict(ours.helper_params, regions=ours.regions)
xi = torch.arange(coords.shape[0], dtype=torch.float32)[:, None]
xp, kw, real = prepare.prepare_input(xi.to(dev), coords.to(dev), batch.to(dev), helper)
xo, kwo, realo = OraclePrepare.prepare_input(xi, coords, batch, dict(ours.helper_params, regions=ours.regions.cpu()))
print("pad rows picking another point:", int((xp.cpu()[:, 0] != xo[:, 0]).sum()), "of", int((~realo).sum()))
print("shift mismatches:", int((kw["combined_shifts"].cpu() != kwo["combined_shifts"]).sum()))
want = ref(x, coords, batch)
g = torch.randn(want.shape, generator=torch.Generator().manual_seed(2))
want.backward(g)
for eng, bwd in ((1, 3), (0, 1), (1, 1), (0, 3)):
    lib.hept_set_engine(eng); lib.hept_set_bwd_variant(bwd)
    ours.zero_grad(set_to_none=True)
    out = ours(x.to(dev), coords.to(dev), batch.to(dev))
    row = (out.detach().cpu() - want.detach()).norm(dim=1) / want.detach().norm(dim=1).clamp_min(1e-12)
    out.backward(g.to(dev))
    errs = {}
    for name in ("W.weight", "attns.0.w_q.weight", "attns.3.w_rpe.weight", "feat_encoder.0.weight", "attns.1.attn.out_linear.bias"):
        a = dict(ours.named_parameters())[name].grad.cpu(); b = dict(ref.named_parameters())[name].grad
        errs[name] = "%.2e" % float((a - b).norm() / b.norm())
    print("engine", eng, "bwd", bwd, "row median %.2e frac>1e-2 %.3f" % (float(row.median()), float((row > 1e-2).float().mean())), errs)

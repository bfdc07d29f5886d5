"""Per-head error of numerators / normalisers against the float64 oracle in the trained-weight regime, ours vs the
reference's fp32 (diagnostic, GPU)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hept_b200 import _lib, ops
from oracle import hept_oracle as O
from tests.helpers import load_case
from tests.test_gpu_parity import oracle_trace, dims_of, to_dev

lib = _lib.load()
dev = torch.device("cuda:0")
res = {}
for name in ("ckpt_l0", "ckpt_l2"):
    cfg, inputs, params, grad_out, gold, meta = load_case(name)
    n = inputs["query"].shape[0]
    d = dims_of(cfg, n)
    di = to_dev(inputs)
    positions = (gold["q_pos"].long(), gold["k_pos"].long())
    t32 = oracle_trace(cfg, inputs, params, torch.float32, positions)
    t64 = oracle_trace(cfg, inputs, params, torch.float64, positions)
    scale = ops.coord_scale(params["w_rpe.weight"].to(dev), d.H, d.D, cfg["num_w_per_dist"])
    print(name, "scale max per head", scale.max(dim=1).values.cpu().tolist())
    pos = torch.stack([gold["q_pos"], gold["k_pos"]]).to(torch.int32).to(dev)
    # block extents in hat space
    sq = O.gather_blocks(t64["q_hat"], positions[0], d.B)
    sk = O.gather_blocks(t64["k_hat"], positions[1], d.B)
    ctr = sk[..., -1:, :]
    print(name, "median |q'|^2 per head", ((sq - ctr) ** 2).sum(-1).median(dim=-1).values.median(dim=-1).values.mean(0).tolist())
    for eng in (0, 1):
        lib.hept_set_engine(eng)
        stage = ops.block_attention_fwd(d, di["query"], di["key"], di["value"], di["coords"], scale, pos)
        numer = stage[..., : d.D].permute(2, 0, 1, 3).cpu().double()
        for h in range(d.H):
            eo = float((numer[:, h] - t64["numer"][:, h]).norm() / t64["numer"][:, h].norm())
            er = float((t32["numer"][:, h].double() - t64["numer"][:, h]).norm() / t64["numer"][:, h].norm())
            res[f"{name}_eng{eng}_h{h}"] = [eo, er]
            print(name, "engine", eng, "head", h, "ours %.2e ref %.2e" % (eo, er))
    lib.hept_set_engine(1)
json.dump(res, open("gpurun_out/regime_b_diag.json", "w"), indent=1)

"""One tracking-60k fwd+bwd through the stage-wise API, twice (warm-up + profiled) — the command ncu wraps.

    ncu --set full --clock-control none --import-source on -k regex:block_attn -s 3 -c 3 -o gpurun_out/prof \
        python tools/profile_step.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from hept_b200 import ops

from hept_b200 import _lib
_lib.load().hept_set_engine(1 if os.environ.get("HEPT_ENGINE", "tcgen05") == "tcgen05" else 0)
_lib.load().hept_set_bwd_variant(int(os.environ.get("HEPT_BWD", "3")))
n_raw = int(sys.argv[1]) if len(sys.argv) > 1 else 60000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
cfg, params, inp, g = bench.make_event(7, n_raw, device="cuda:0")
dev = torch.device("cuda:0")
inp = {k: v.to(dev) for k, v in inp.items()}
n = inp["query"].shape[0]
d = ops.Dims(N=n, H=cfg["num_heads"], D=cfg["h_dim"], C=cfg["coords_dim"], T=cfg["n_hashes"], B=cfg["block_size"], raw_size=n)
w, al = params["w_rpe.weight"].to(dev), params["e2lsh.alpha"].to(dev)
wo, bo = params["out_linear.weight"].to(dev), params["out_linear.bias"].to(dev)
gout = torch.randn(n, d.D, device=dev)
for _ in range(reps):
    out, den, scale, pos = ops.attention_fwd(d, inp["query"], inp["key"], inp["value"], inp["coords"], w,
                                             cfg["num_w_per_dist"], al, combined_shifts=inp["combined_shifts"])
    y = ops.out_linear_fwd(d, out, wo, bo)
    gpre, dwo, dbo = ops.out_linear_bwd(d, gout, wo, out)
    dq, dk, dv, dscale = ops.attention_bwd(d, inp["query"], inp["key"], inp["value"], inp["coords"], scale, pos, out, den, gpre)
    dw = ops.coord_scale_backward(w, scale, dscale, d.H, d.D, cfg["num_w_per_dist"])
torch.cuda.synchronize()
print("done", float(out.abs().mean()), float(dq.abs().mean()))

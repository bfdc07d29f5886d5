"""One tracking-60k fwd+bwd at the Attn-block boundary (norm1 + w_q/w_k/w_v front, attention, out_linear) through the
stage-wise C ABI, `reps` times — the command ncu wraps for the round-2 launch lists.

    ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv \
        python tools/profile_block.py [n_raw] [reps]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from hept_b200 import ops

n_raw = int(sys.argv[1]) if len(sys.argv) > 1 else 60000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda:0")
cfg, params, inp, g = bench.make_event(7, n_raw, device=dev)
inp = {k: v.to(dev) for k, v in inp.items()}
n = inp["coords"].shape[0]
H, D = cfg["num_heads"], cfg["h_dim"]
d = ops.Dims(N=n, H=H, D=D, C=cfg["coords_dim"], T=cfg["n_hashes"], B=cfg["block_size"], raw_size=n)
w, al = params["w_rpe.weight"].to(dev), params["e2lsh.alpha"].to(dev)
wo, bo = params["out_linear.weight"].to(dev), params["out_linear.bias"].to(dev)
gen = torch.Generator().manual_seed(3)
x = (torch.randn(n, D, generator=gen) * 0.7).to(dev)
gam, bet = torch.ones(D, device=dev), torch.zeros(D, device=dev)
wq, wk, wv = ((torch.randn(H * D, D, generator=gen) / D ** 0.5).to(dev) for _ in range(3))
sh32 = inp["combined_shifts"].to(torch.int32)
gout = torch.randn(n, D, device=dev)
for _ in range(reps):
    q, k, v, xn, wt = ops.attn_qkv_fwd(x, gam, bet, wq, wk, wv, H, D, 1e-5)
    out, den, scale, pos = ops.attention_fwd(d, q, k, v, inp["coords"], w, cfg["num_w_per_dist"], al, combined_shifts=sh32)
    y = ops.out_linear_fwd(d, out, wo, bo)
    gpre, dwo, dbo = ops.out_linear_bwd(d, gout, wo, out)
    dq, dk, dv, dscale = ops.attention_bwd(d, q, k, v, inp["coords"], scale, pos, out, den, gpre)
    dw = ops.coord_scale_backward(w, scale, dscale, d.H, d.D, cfg["num_w_per_dist"])
    dx, dgam, dbet, dwq, dwk, dwv = ops.attn_qkv_bwd(x, xn, gam, wt, dq, dk, dv, H, D, 1e-5)
torch.cuda.synchronize()
print("done", float(y.abs().mean()), float(dx.abs().mean()))

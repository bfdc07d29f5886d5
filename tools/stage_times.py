"""CUDA-event time of every native stage of one tracking-60k fwd+bwd (inputs rotate over 4 events, far beyond L2).

    python tools/stage_times.py [n_hits]     -> gpurun_out/stage_times.json
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench
from hept_b200 import _lib, ops

n_raw = int(sys.argv[1]) if len(sys.argv) > 1 else 60000
lib = _lib.load()
dev = torch.device("cuda:0")
sets = []
for i in range(4):
    cfg, params, inp, g = bench.make_event(7 + i, n_raw, device="cuda:0")
    inp = {k: v.to(dev) for k, v in inp.items()}
    sets.append(inp)
n = sets[0]["query"].shape[0]
d = ops.Dims(N=n, H=cfg["num_heads"], D=cfg["h_dim"], C=cfg["coords_dim"], T=cfg["n_hashes"], B=cfg["block_size"], raw_size=n)
K = cfg["num_w_per_dist"]
w, al = params["w_rpe.weight"].to(dev), params["e2lsh.alpha"].to(dev)
scale = ops.coord_scale(w, d.H, d.D, K)
mid = []
for s in sets:
    proj, span = ops.hash_project(d, s["query"], s["key"], s["coords"], scale, al)
    keys = ops.keys_from_packed_shifts(d, proj, span, s["combined_shifts"])
    pos = ops.segmented_argsort(keys)
    stage = ops.block_attention_fwd(d, s["query"], s["key"], s["value"], s["coords"], scale, pos)
    out, den = ops.or_combine(d, stage)
    mid.append(dict(proj=proj, span=span, keys=keys, pos=pos, stage=stage, out=out, den=den, g=torch.randn_like(out)))


def ev(fn, reps=12):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


S = lambda i: sets[i % 4]
M = lambda i: mid[i % 4]
t = {}
t["hash_project"] = ev(lambda i: ops.hash_project(d, S(i)["query"], S(i)["key"], S(i)["coords"], scale, al))
t["keys"] = ev(lambda i: ops.keys_from_packed_shifts(d, M(i)["proj"], M(i)["span"], S(i)["combined_shifts"]))
t["argsort"] = ev(lambda i: ops.segmented_argsort(M(i)["keys"]))
t["hat+tiles_fwd"] = ev(lambda i: ops.block_attention_fwd(d, S(i)["query"], S(i)["key"], S(i)["value"], S(i)["coords"], scale, M(i)["pos"]))
t["or_combine"] = ev(lambda i: ops.or_combine(d, M(i)["stage"]))
t["fwd_call"] = ev(lambda i: ops.attention_fwd(d, S(i)["query"], S(i)["key"], S(i)["value"], S(i)["coords"], w, K, al,
                                               combined_shifts=S(i)["combined_shifts"]))
t["bwd_call"] = ev(lambda i: ops.attention_bwd(d, S(i)["query"], S(i)["key"], S(i)["value"], S(i)["coords"], scale,
                                               M(i)["pos"], M(i)["out"], M(i)["den"], M(i)["g"]))
wo, bo = params["out_linear.weight"].to(dev), params["out_linear.bias"].to(dev)
go = torch.randn(n, d.D, device=dev)
t["out_linear_fwd"] = ev(lambda i: ops.out_linear_fwd(d, M(i)["out"], wo, bo))
t["out_linear_bwd"] = ev(lambda i: ops.out_linear_bwd(d, go, wo, M(i)["out"]))
t["out_linear_bwd_params_only"] = ev(lambda i: ops.out_linear_bwd(d, go, wo, M(i)["out"], need_input_grad=False))
lib.hept_set_bwd_stage_mask(3)
t["bwd_pre+tiles"] = ev(lambda i: ops.attention_bwd(d, S(i)["query"], S(i)["key"], S(i)["value"], S(i)["coords"], scale,
                                                    M(i)["pos"], M(i)["out"], M(i)["den"], M(i)["g"]))
lib.hept_set_bwd_stage_mask(7)
# the rows either side of the attention call: prepare_input (a13-a17) and the Attn front
from hept_b200 import prepare, synthetic

craw, batch = synthetic.batched_cloud([n_raw], cfg["coords_dim"], 7)
craw, batch = craw.to(dev), batch.to(dev)
helper = {"block_size": cfg["block_size"], "regions": params["regions"].to(dev), "num_heads": cfg["num_heads"]}
xin = torch.zeros(n_raw, 1, device=dev)
t["prepare_input_batched"] = ev(lambda i: prepare.prepare_input(xin, craw, batch, helper, sizes=[n_raw]))
t["prepare_input_single"] = ev(lambda i: prepare.prepare_input_single(xin, craw, helper))
if ops.attn_qkv_supported(d.H, d.D):
    gen = torch.Generator().manual_seed(3)
    x = (torch.randn(n, d.D, generator=gen) * 0.7).to(dev)
    gam, bet = torch.ones(d.D, device=dev), torch.zeros(d.D, device=dev)
    wq, wk, wv = ((torch.randn(d.H * d.D, d.D, generator=gen) / d.D ** 0.5).to(dev) for _ in range(3))
    q_, k_, v_, xn, wt = ops.attn_qkv_fwd(x, gam, bet, wq, wk, wv, d.H, d.D, 1e-5)
    t["attn_qkv_fwd"] = ev(lambda i: ops.attn_qkv_fwd(x, gam, bet, wq, wk, wv, d.H, d.D, 1e-5))
    t["attn_qkv_bwd"] = ev(lambda i: ops.attn_qkv_bwd(x, xn, gam, wt, S(i)["query"], S(i)["key"], S(i)["value"], d.H, d.D, 1e-5))
# loss and metrics at this size
from hept_b200 import metrics as hmetrics
from hept_b200.losses import InfoNCELoss

cidt, recons, pts, pairs = (a.to(dev) for a in synthetic.tracking_truth(n_raw, 5))
emb = (torch.randn(n_raw, 12, generator=torch.Generator().manual_seed(9)) * 0.5).to(dev).requires_grad_(True)
crit = InfoNCELoss(0.05, "l2_rbf")
t["infonce_fwd"] = ev(lambda i: crit(emb, pairs, cidt, recons, pts))


def _fb(i):
    emb.grad = None
    crit(emb, pairs, cidt, recons, pts).backward()


t["infonce_fwd_bwd"] = ev(_fb)
t["infonce_pairs"] = float(pairs.shape[1])
mask = hmetrics.point_filter(cidt, recons, pts, 0.9)
t["knn_metrics"] = ev(lambda i: hmetrics.acc_and_pr_at_k(emb.detach(), cidt, mask, "l2_rbf", K=31), reps=3)   # synthetic particles have up to ~25 hits
t["knn_queries"] = float(mask.sum())
t = {k: round(v, 1) for k, v in t.items()}
print(json.dumps(t))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(t, open(os.path.join(ROOT, "gpurun_out", "stage_times.json"), "w"))

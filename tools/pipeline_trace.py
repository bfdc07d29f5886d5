"""Timeline of the warp-specialised tile kernels: clock64 stamps of every hand-off of CTA 0 (trace.cuh), first 64 tiles.

    make -C hept_b200/csrc TRACE=1 && python tools/pipeline_trace.py

Prints, per tile, each event's offset from the first stamp and the steady-state period / phases; writes
gpurun_out/pipeline_trace.json.  The stamps cost the stamping warp ~100 cycles each: the traced CTA runs 10-15 % slower than
the others (compare its period with the per-CTA durations), and phases that hold several stamps are stretched -- per-line
stall samples (tools/ncu_lines.py) are the cross-check."""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("HEPT_LIB", os.path.join(ROOT, "hept_b200", "libhept_sm100_trace.so"))
import torch

import bench
from hept_b200 import _lib, ops

FWD = ["E_SREADY", "E_PREADY", "E_ODONE", "E_OUT", "P_QKFREE", "P_QKFULL", "P_VFREE", "P_VFULL", "P_ISSUED",
       "M_S_GO", "M_S_ISSUED", "M_PV_GO", "M_PV_ISSUED"]
BWD = ["E_QREADY", "E_DSRDY", "E_KREADY", "E_PTRDY", "E_DQDONE", "E_DQOUT", "E_DVDONE", "E_DSTRDY", "E_DVOUT", "E_DKDONE",
       "E_END", "P_KFREE", "P_KFULL", "P_ISSUED", "P_MFREE", "P_MFULL", "M_DQ_GO", "M_DV_GO", "M_DK_GO", "M_SQ_GO",
       "M_SK_GO", "M_END", "E_DSTISS", "E_DVWAIT", "E_DQACC", "E_DKACC"]
TILES, EVENTS = 64, 32

lib = _lib.load()
lib.hept_set_engine(1)
lib.hept_set_bwd_variant(3)
cfg, params, inp, g = bench.make_event(7, 60000, device="cuda:0")
dev = torch.device("cuda:0")
inp = {k: v.to(dev) for k, v in inp.items()}
n = inp["query"].shape[0]
d = ops.Dims(N=n, H=cfg["num_heads"], D=cfg["h_dim"], C=cfg["coords_dim"], T=cfg["n_hashes"], B=cfg["block_size"], raw_size=n)
w, al = params["w_rpe.weight"].to(dev), params["e2lsh.alpha"].to(dev)
gpre = torch.randn(n, d.H * d.D, device=dev)
tf = torch.zeros(EVENTS * TILES + 4 * 1024, dtype=torch.int64, device=dev)
tb = torch.zeros(EVENTS * TILES + 4 * 1024, dtype=torch.int64, device=dev)
for fn, buf in (("hept_debug_trace_fwd", tf), ("hept_debug_trace_bwd", tb)):
    f = getattr(lib, fn)       # only the TRACE build exports these
    f.argtypes, f.restype = [ctypes.c_void_p], ctypes.c_int
    assert f(ctypes.c_void_p(buf.data_ptr())) == 0
for _ in range(3):
    tf.zero_(); tb.zero_()
    out, den, scale, pos = ops.attention_fwd(d, inp["query"], inp["key"], inp["value"], inp["coords"], w,
                                             cfg["num_w_per_dist"], al, combined_shifts=inp["combined_shifts"])
    ops.attention_bwd(d, inp["query"], inp["key"], inp["value"], inp["coords"], scale, pos, out, den, gpre)
torch.cuda.synchronize()
res = {}
for name, ev, buf in (("fwd", FWD, tf), ("bwd", BWD, tb)):
    t = buf.cpu()[: EVENTS * TILES].view(EVENTS, TILES)[: len(ev)]
    t0 = int(t[t > 0].min())
    rel = (t - t0).clamp_min(-1)
    print(f"==== {name}: cycles since the first stamp; rows = tiles of CTA 0")
    print("tile " + " ".join(f"{e:>10}" for e in ev))
    for it in range(8, 20):
        print(f"{it:4d} " + " ".join(f"{int(rel[e, it]):>10}" for e in range(len(ev))))
    lo, hi = 10, 60
    period = float(t[0, hi] - t[0, lo]) / (hi - lo)
    phase = {e: float((t[i, lo:hi] - t[0, lo:hi]).double().mean()) for i, e in enumerate(ev)}
    print(f"period {period:.0f} cycles/tile; mean offset from {ev[0]}: " + ", ".join(f"{k}={v:.0f}" for k, v in phase.items()))
    res[name] = {"period_cycles": period, "phase": phase, "events": ev, "stamps": rel[:, :40].tolist()}
per_sm = {}
for name, buf in (("fwd", tf), ("bwd", tb)):
    cta = buf.cpu()[EVENTS * TILES:].view(-1, 4)
    cta = cta[cta[:, 1] > 0]
    dur = (cta[:, 1] - cta[:, 0]).double()
    print(f"{name} per-CTA duration (cycles): n={len(cta)} min={dur.min():.0f} median={dur.median():.0f} max={dur.max():.0f}")
    res[name + "_cta_cycles"] = dur.tolist()
    res[name + "_cta_smid"] = cta[:, 2].tolist()
    per_sm[name] = {int(sm): float(d) for sm, d in zip(cta[:, 2], dur)}
common = sorted(set(per_sm["fwd"]) & set(per_sm["bwd"]))
f = torch.tensor([per_sm["fwd"][k] for k in common]); b = torch.tensor([per_sm["bwd"][k] for k in common])
print("fwd/bwd per-SM duration correlation:", float(torch.corrcoef(torch.stack([f, b]))[0, 1]))
order = sorted(common, key=lambda k: per_sm["bwd"][k])
print("fastest SMs (bwd):", order[:12], " slowest:", order[-12:])
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "pipeline_trace.json"), "w"))

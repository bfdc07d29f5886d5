"""The incumbent on the same GPU: the reference's own formulation (ATen library kernels: bmm, argsort, gather,
einsum, exp ... — here through the oracle, which restates it op for op) run on the B200 in eager fp32, next to the
hand-written path.  BASELINE.md section 4 asks for this number; it is recorded in gpurun_out/incumbent.json."""
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _time(fn, steps=5, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def test_hand_written_path_beats_the_library_formulation_on_the_same_gpu():
    import bench
    from hept_b200 import HEPTAttention
    from oracle import hept_oracle as O

    torch.backends.cuda.matmul.allow_tf32 = False
    dev = torch.device("cuda:0")
    cfg, params, inputs, g = bench.make_event(5, 60000)
    di = {k: v.to(dev) for k, v in inputs.items()}
    pd = {k: v.to(dev) for k, v in params.items()}
    gd = g.to(dev)
    ms_ref = _time(lambda: O.forward_backward(di, pd, cfg, gd, torch.float32), steps=3, warmup=1)

    mod = HEPTAttention(30, **cfg)
    mod.load_state_dict({k: params[k] for k in ("out_linear.weight", "out_linear.bias", "e2lsh.alpha")}, strict=True)
    mod = mod.to(dev)
    w_rpe = torch.nn.Linear(50, 192)
    w_rpe.load_state_dict({"weight": params["w_rpe.weight"], "bias": params["w_rpe.bias"]})
    w_rpe = w_rpe.to(dev)
    q, k, v = (di[x].clone().requires_grad_(True) for x in ("query", "key", "value"))

    def ours():
        for t in (q, k, v):
            t.grad = None
        mod(q, k, v, w_rpe=w_rpe, coords=di["coords"], combined_shifts=di["combined_shifts"]).backward(gd)

    ms_ours = _time(ours, steps=10, warmup=3)
    peak_ref = torch.cuda.max_memory_allocated() / 1e9
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/incumbent.json", "w") as f:
        json.dump({"workload": "HEPTAttention fwd+bwd, 60000 hits, fp32 eager", "library_formulation_ms": ms_ref,
                   "hept_b200_ms": ms_ours, "speedup": ms_ref / ms_ours, "peak_mem_gb_incl_reference": peak_ref}, f, indent=1)
    assert ms_ours < ms_ref

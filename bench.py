"""bench.py — HEPT attention fwd+bwd throughput on synthetic tracking-60k-shaped events.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one event: ``HEPTAttention`` forward + backward w.r.t.
q, k, v, w_rpe.weight and out_linear.* (BASELINE.json configs[1]: ~60k hits, block_size 100, 3 hash
tables, 8 heads x 24 dims, coords_dim 6).  Prints ONE JSON line (rank 0).

  value      hits/s, whole job, inputs already resident in HBM, CUDA-event timed, max over ranks
  e2e        the same metric with HOST (pinned) inputs, H2D of the inputs and D2H of the output + parameter gradients
             inside the timed region, at the Attn-block boundary: the 96-byte activation row of a hit
             crosses the link and norm1 + w_q / w_k / w_v run in the library in front of the same HEPTAttention
  e2e_module_boundary   the same at the reference module's own boundary (q, k, v cross the link: PCIe-bound)
  roofline   dominant kernel, algorithmic bytes per launch / its CUDA-event duration, vs the measured HBM peak
  cpu_baseline  the oracle (a torch-CPU restatement of the reference, kind "port") timed on this box's cores

``--impl reference`` times that CPU path alone with the same metric / config (rank 0 only).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "HEPT attention hits/sec fwd+bwd, tracking-60k shape; % of HBM roofline"
UNIT = "hits/s"
N_RAW = 60000
FWD_BWD_BYTES_PER_HIT = 8336          # compulsory bytes at the module boundary, fp32 I/O


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


WORKLOAD = "tracking-60k fwd+bwd (HEPTAttention, H=8 D=24 C=6 T=3 B=100), one 60000-hit event per step per GPU"


def make_event(seed: int, n_raw: int = N_RAW, device=None):
    """Synthetic tracking-shaped event -> CPU tensors.  ``device`` given: the product's prepare_input (CUDA kernels behind
    the C ABI) builds the hash codes; ``device=None`` (the CPU reference arm only): the oracle's restatement does."""
    from hept_b200 import synthetic

    cfg = dict(synthetic.TRACKING)
    coords_raw, batch = synthetic.batched_cloud([n_raw], cfg["coords_dim"], seed)
    params = synthetic.module_params(cfg, 0)
    if device is not None:
        from hept_b200 import prepare

        helper = {"block_size": cfg["block_size"], "regions": params["regions"].to(device), "num_heads": cfg["num_heads"]}
        _, kw, _ = prepare.prepare_input(torch.zeros(n_raw, 1, device=device), coords_raw.to(device), batch.to(device), helper,
                                         sizes=[n_raw])
        kw = {k: v.cpu() for k, v in kw.items()}
    else:
        from oracle import hept_oracle as O

        _, kw, _ = O.prepare_batched(torch.zeros(n_raw, 1), coords_raw, batch, params["regions"], cfg["block_size"],
                                     cfg["num_heads"])
    n = kw["coords"].shape[0]
    q, k, v = synthetic.qkv(n, cfg, seed)
    g = torch.randn(n, cfg["h_dim"], generator=torch.Generator().manual_seed(seed + 5))
    return cfg, params, dict(query=q, key=k, value=v, coords=kw["coords"], combined_shifts=kw["combined_shifts"]), g


# ------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz, self._stop = [], set(), None, threading.Event()
        self._thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nv = None

    def _run(self):
        nv = self._nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.02)

    def __enter__(self):
        if self._nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        s = sorted(self.samples)
        med = s[len(s) // 2] if s else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


# --------------------------------------------------------------------------------------- CPU baseline
def cpu_port_rate(n_hits: int, steps: int, warmup: int):
    """Oracle (torch-CPU restatement of the reference) fwd+bwd on ``n_hits``-hit events -> hits/s, seconds/step."""
    from oracle import hept_oracle as O

    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    cfg, params, inputs, g = make_event(1234, n_hits)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        O.forward_backward(inputs, params, cfg, g, torch.float32)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    dt = sum(times) / len(times)
    return n_hits / dt, dt, threads


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port; the reference is Python and cannot travel)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # bounded sample: events small enough that warmup+steps finish in ~2 minutes (cost is linear in hits)
    probe_rate, _, threads = cpu_port_rate(6000, 1, 1)
    budget_s = 110.0
    n = int(min(N_RAW, max(2000, probe_rate * budget_s / max(1, args.steps + args.warmup))) // 100 * 100)
    rate, dt, threads = cpu_port_rate(n, args.steps, args.warmup)
    line = {
        "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "impl": "reference",
        "config": {"workload": WORKLOAD},
        "sample": f"{n}-hit events, cost linear in hits",
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{args.steps} fwd+bwd steps on {n}-hit synthetic tracking events, torch CPU eager fp32"},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch.distributed as dist

    from hept_b200 import HEPTAttention, ops, sharding, _lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    try:   # pin this rank to the CPUs next to its GPU before any pinned host buffer exists: with 8 ranks the H2D copies of the
        import pynvml   # e2e leg otherwise cross the socket interconnect

        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
    except Exception:
        pass
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cuda.matmul.allow_tf32 = False
    if args.engine is not None:
        _lib.load().hept_set_engine(1 if args.engine == "tcgen05" else 0)
    if args.bwd is not None:
        _lib.load().hept_set_bwd_variant(args.bwd)
    engine_name = ("tcgen05" if _lib.load().hept_get_engine() else "simt") + f"+bwd{_lib.load().hept_get_bwd_variant()}"

    n_sets = 4                                  # rotate over 4 events: ~0.75 GB of inputs, far beyond the 126 MB L2
    events = [make_event(100 * rank + i, device=dev) for i in range(n_sets)]
    cfg, params = events[0][0], events[0][1]
    mod = HEPTAttention(cfg["h_dim"] + cfg["coords_dim"], **cfg)
    mod.load_state_dict({k: params[k] for k in ("out_linear.weight", "out_linear.bias", "e2lsh.alpha")}, strict=True)
    mod = mod.to(dev)
    w_rpe = torch.nn.Linear(params["w_rpe.weight"].shape[1], params["w_rpe.weight"].shape[0])
    w_rpe.load_state_dict({"weight": params["w_rpe.weight"], "bias": params["w_rpe.bias"]})
    w_rpe = w_rpe.to(dev)
    trainable = [w_rpe.weight, mod.out_linear.weight, mod.out_linear.bias]
    # gradients live in one persistent flat buffer (hept_b200/sharding.py): ONE NCCL launch per step for the collective
    bucket = sharding.GradBucket(trainable, p2p=not args.nccl)

    host = [{k: v.pin_memory() for k, v in e[2].items()} for e in events]
    gouts = [e[3].to(dev) for e in events]
    resident = [{k: v.to(dev) for k, v in h.items()} for h in host]
    for r in resident:
        for k in ("query", "key", "value"):
            r[k].requires_grad_(True)
    n_hits = resident[0]["query"].shape[0]

    def step(inp, g):
        bucket.zero()
        for k in ("query", "key", "value"):
            inp[k].grad = None
        out = mod(inp["query"], inp["key"], inp["value"], w_rpe=w_rpe, coords=inp["coords"],
                  combined_shifts=inp["combined_shifts"])
        out.backward(g)
        if world > 1:
            bucket.allreduce()
        return out

    def timed(fn, steps, warmup):
        for i in range(warmup):
            fn(i)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ops.launch_count(reset=True)
        e0.record()
        for i in range(steps):
            fn(warmup + i)
        e1.record()
        torch.cuda.synchronize()
        launches = ops.launch_count(reset=True)
        if world > 1:
            dist.barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches

    # ---- device-resident throughput -----------------------------------------------------------------
    with ClockSampler(local) as clocks:
        ms, launches = timed(lambda i: step(resident[i % n_sets], gouts[i % n_sets]), args.steps, args.warmup)
    value = world * args.steps * N_RAW / (ms * 1e-3)

    # ---- end to end with HOST buffers --------------------------------------------------------------
    # Double-buffered: the H2D copy of step i+1 runs on a copy stream while step i computes; every step still
    # copies its own inputs from pinned host memory and reads its results back inside the timed region.
    def e2e_leg(host_sets, grad_keys, run, grads, out_width, do_h2d=True, do_d2h=True):   # grads: the GradBucket that is read back
        staging = [{k: torch.empty_like(v, device=dev) for k, v in host_sets[0].items()} for _ in range(2)]
        out_host = torch.empty(n_hits, out_width).pin_memory()
        grad_host = torch.empty_like(grads.flat, device="cpu").pin_memory()
        grad_stage = torch.empty_like(grads.flat)                       # device copy the read-back stream reads from
        h2d = sum(v.numel() * v.element_size() for v in host_sets[0].values())
        d2h = out_host.numel() * 4 + grad_host.numel() * 4
        copy_stream = torch.cuda.Stream(device=dev)
        back_stream = torch.cuda.Stream(device=dev)      # read-back of step i overlaps the compute of step i + 1
        ready = [torch.cuda.Event() for _ in range(2)]
        consumed = [torch.cuda.Event() for _ in range(2)]
        computed, read_back = torch.cuda.Event(), torch.cuda.Event()
        issued = set()

        def prefetch(i):
            if i in issued:
                return
            issued.add(i)
            b = i % 2
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[b])          # buffer b was last read by step i-2
                if do_h2d or i < 2:
                    for k, v in host_sets[i % n_sets].items():
                        staging[b][k].copy_(v, non_blocking=True)
                ready[b].record(copy_stream)

        def e2e_step(i):
            prefetch(i)
            prefetch(i + 1)
            b = i % 2
            cur = torch.cuda.current_stream()
            cur.wait_event(ready[b])
            inp = dict(staging[b])
            for k in grad_keys:
                inp[k] = inp[k].detach().requires_grad_(True)
            out = run(inp, gouts[i % n_sets])
            consumed[b].record(cur)
            cur.wait_event(read_back)                        # the previous step's gradients have left their staging copies
            grad_stage.copy_(grads.flat)
            computed.record(cur)
            out = out.detach()
            out.record_stream(back_stream)
            with torch.cuda.stream(back_stream):
                back_stream.wait_event(computed)
                if do_d2h:
                    out_host.copy_(out, non_blocking=True)
                    grad_host.copy_(grad_stage, non_blocking=True)
                read_back.record(back_stream)

        steps = max(3, min(args.steps, 10))
        ms_leg, _ = timed(e2e_step, steps, 3)
        return {"value": world * steps * N_RAW / (ms_leg * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": ms_leg / steps}

    # (1) the module boundary of the reference, HEPTAttention.forward(query, key, value, ...): q, k, v cross the link
    e2e_module = e2e_leg(host, ("query", "key", "value"), step, bucket, cfg["h_dim"])
    e2e_module["boundary"] = ("HEPTAttention.forward: q, k, v (N,192) fp32, coords, int64 codes from pinned host memory; "
                              "output (N,24) + parameter gradients read back")

    # (2) the Attn-block boundary: the activation row x (N,24) crosses the link, norm1 + w_q / w_k / w_v run
    # in the library (csrc/attn_block.cu) in front of the same HEPTAttention; MORE work per hit than (1), 12x fewer bytes
    from hept_b200.attention import attn_front

    g0 = torch.Generator().manual_seed(4242 + rank)
    norm1 = torch.nn.LayerNorm(cfg["h_dim"]).to(dev)
    w_qkv = [torch.nn.Linear(cfg["h_dim"], cfg["num_heads"] * cfg["h_dim"], bias=False).to(dev) for _ in range(3)]
    front_params = [norm1.weight, norm1.bias] + [w.weight for w in w_qkv]
    bucket_blk = sharding.GradBucket(trainable + front_params, p2p=not args.nccl)   # re-homes the three attention parameters' .grad too
    host_x = [{"x": (torch.randn(n_hits, cfg["h_dim"], generator=g0) * 0.7).pin_memory(), "coords": h["coords"],
               "combined_shifts32": h["combined_shifts"].to(torch.int32).pin_memory()} for h in host]

    def block_step(inp, g):
        bucket_blk.zero()
        q, k, v = attn_front(inp["x"], norm1, w_qkv[0], w_qkv[1], w_qkv[2], cfg["num_heads"])
        out = mod(q, k, v, w_rpe=w_rpe, coords=inp["coords"], combined_shifts32=inp["combined_shifts32"])
        out.backward(g)
        if world > 1:
            bucket_blk.allreduce()
        return out

    resident_x = [{k: v.to(dev) for k, v in h.items()} for h in host_x]
    for r in resident_x:
        r["x"].requires_grad_(True)

    def block_resident(i):
        r = resident_x[i % n_sets]
        r["x"].grad = None
        block_step(r, gouts[i % n_sets])

    ms_blk, launches_blk = timed(block_resident, args.steps, args.warmup)
    e2e_block = e2e_leg(host_x, ("x",), block_step, bucket_blk, cfg["h_dim"])
    e2e_block["boundary"] = ("Attn block front + HEPTAttention: x (N,24) fp32, coords, int32 codes from pinned host memory; norm1 and "
                             "w_q / w_k / w_v computed on the device; output (N,24) + parameter gradients read back")
    # the same leg with the compute captured in CUDA graphs (hept_b200/graphed.py): two graph launches per step instead of ~45
    # kernel launches — on a box where eight ranks share the host's PCIe / launch path, the launches, not the 13 MB of copies,
    # are what the concurrent DMA traffic delays
    from hept_b200.graphed import graphed_attn_block

    gstep = graphed_attn_block(mod, w_rpe, norm1, w_qkv[0], w_qkv[1], w_qkv[2], resident_x[0]["x"].detach().requires_grad_(True),
                               resident_x[0]["coords"], resident_x[0]["combined_shifts32"])

    def block_step_graphed(inp, g):
        bucket_blk.zero()
        out = gstep(inp["x"], inp["coords"], inp["combined_shifts32"])
        out.backward(g)
        if world > 1:
            bucket_blk.allreduce()
        return out

    e2e_graph = e2e_leg(host_x, ("x",), block_step_graphed, bucket_blk, cfg["h_dim"])
    e2e_block["cuda_graph_replay"] = {"value": e2e_graph["value"], "ms_per_step": e2e_graph["ms_per_step"]}
    if args.e2e_diag:      # where does the end-to-end leg lose time (multi-GPU boxes): the same leg without one copy direction
        e2e_block["diag_no_d2h_ms"] = e2e_leg(host_x, ("x",), block_step, bucket_blk, cfg["h_dim"], do_d2h=False)["ms_per_step"]
        e2e_block["diag_no_h2d_ms"] = e2e_leg(host_x, ("x",), block_step, bucket_blk, cfg["h_dim"], do_h2d=False)["ms_per_step"]
        e2e_block["diag_no_copies_ms"] = e2e_leg(host_x, ("x",), block_step, bucket_blk, cfg["h_dim"], do_h2d=False, do_d2h=False)["ms_per_step"]
    e2e_block["device_resident_same_boundary"] = {"value": world * args.steps * N_RAW / (ms_blk * 1e-3),
                                                  "ms_per_step": ms_blk / args.steps, "gpu_launches": launches_blk}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD},
        "details": {"l2": f"inputs rotate over {n_sets} events (~190 MB each) so no step re-reads L2-resident inputs",
                    "collective": ("none" if world == 1 else "one-shot all-reduce of the flat gradient bucket over NVLink peer memory "
                                   "(hept_p2p_allreduce)" if bucket.uses_p2p else f"NCCL all-reduce (AVG) of the flat gradient bucket ({bucket.p2p_error})"),
                    "tile_engine": engine_name},
        "e2e": e2e_block,
        "e2e_module_boundary": e2e_module,
        "gpu_launches": launches,
        "clocks": clocks.summary(),
    }

    # ---- BASELINE.json configs[4]: the tracking training step sharded by event -----------------------------------
    # 4-layer HEPT Transformer (329 364 parameters) forward + InfoNCE loss (l2_rbf, tau 0.05: tracking_trans_hept.yaml:22-25;
    # hept_b200/losses.py on the library's kernels) + backward + Adam, one 60 000-hit event per rank per step, the gradients of
    # all parameters in ONE flat bucket all-reduced (averaged) with one NCCL launch.  Point pairs / particle ids are synthetic.
    if not args.no_train:
        from hept_b200 import synthetic
        from hept_b200.model import Transformer

        tcfg = {k: v for k, v in synthetic.TRACKING.items() if k != "coords_dim"}
        torch.manual_seed(1234)                         # the same initial weights on every rank
        model = Transformer(in_dim=15, coords_dim=6, **tcfg).to(dev)
        tbucket = sharding.GradBucket(model.parameters(), p2p=not args.nccl)
        opt = torch.optim.Adam(tbucket.params, lr=1e-3, fused=True)
        tcoords = [synthetic.point_cloud(N_RAW, 6, 1000 + 10 * rank + i).to(dev) for i in range(2)]
        tx = [(torch.randn(N_RAW, 15, generator=torch.Generator().manual_seed(rank + 7 * i)) * 0.5).to(dev) for i in range(2)]
        tbatch = torch.zeros(N_RAW, dtype=torch.long, device=dev)
        from hept_b200.losses import InfoNCELoss

        crit = InfoNCELoss(tau=0.05, dist_metric="l2_rbf")
        truth = [tuple(t.to(dev) for t in synthetic.tracking_truth(N_RAW, 10 * rank + i)) for i in range(2)]

        def train_step(i):
            tbucket.zero()
            out = model(tx[i % 2], tcoords[i % 2], tbatch)
            cid, recons, pts, pairs = truth[i % 2]
            crit(out, pairs, cid, recons, pts).backward()
            if world > 1:
                tbucket.allreduce()
            opt.step()

        t_steps = max(3, min(args.steps, 6))
        ms_t, launches_t = timed(train_step, t_steps, 3)
        line["train_step"] = {"workload": "tracking Transformer (4 HEPT layers) fwd + bwd + Adam, one 60000-hit event per rank per step",
                              "value": world * t_steps * N_RAW / (ms_t * 1e-3), "unit": UNIT, "ms_per_step": ms_t / t_steps,
                              "allreduce_bytes": tbucket.nbytes if world > 1 else 0, "collective": "hept_p2p_allreduce (one kernel over NVLink peer memory)" if tbucket.uses_p2p else "one NCCL all-reduce (AVG) of the flat gradient bucket",
                              "native_launches_per_step": launches_t / t_steps,
                              "loss": f"InfoNCE (l2_rbf, tau 0.05) over {truth[0][3].shape[1]} synthetic point pairs, on the library's kernels"}

    if rank == 0:
        peak, peak_src = load_peaks()
        line["path_roofline"] = {"bytes_per_hit": FWD_BWD_BYTES_PER_HIT,
                                 "achieved_gbs": FWD_BWD_BYTES_PER_HIT * value / world / 1e9,
                                 "frac": FWD_BWD_BYTES_PER_HIT * value / world / 1e9 / peak}
        line["roofline"] = kernel_roofline(resident, gouts, cfg, mod, w_rpe, peak, peak_src, n_sets, value / world)
        inc, ipath = _committed(PROFILE_TAG + "_incumbent.json")
        if inc:      # the reference's formulation on the same GPU (library kernels), measured by tests/test_gpu_incumbent.py
            line["incumbent_gpu"] = dict(inc, source=ipath)
        if world == 1:
            rate, dt, threads = cpu_port_rate(N_RAW, 2, 1)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": "2 fwd+bwd steps (after 1 warm-up) on one 60000-hit event, torch CPU eager fp32"}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


FWD_BYTES_PER_HIT, BWD_BYTES_PER_HIT = 2616, 5720      # compulsory bytes of the forward / backward call
FLOP_PER_HIT = 952128                                    # algorithmic fp32 FLOPs, fwd + bwd
PROFILE_TAG = "r2"                                       # profiles/<tag>_traffic.json, <tag>_ncu_*.csv, <tag>_incumbent.json


def _committed(name):
    path = os.path.join(ROOT, "profiles", name)
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "profiles/" + name
    return None, None


def kernel_roofline(resident, gouts, cfg, mod, w_rpe, peak, peak_src, n_sets, path_hits_per_s):
    """The dominant kernel (the tcgen05 backward tile kernel) against the HBM roofline with SURVEY.md 8(d)'s bytes, timed live
    with CUDA events, next to what bounds it according to the committed ncu captures."""
    from hept_b200 import ops, _lib

    lib = _lib.load()
    H, D, C, T, B = cfg["num_heads"], cfg["h_dim"], cfg["coords_dim"], cfg["n_hashes"], cfg["block_size"]
    n = resident[0]["query"].shape[0]
    d = ops.Dims(N=n, H=H, D=D, C=C, T=T, B=B, raw_size=n)
    K = cfg["num_w_per_dist"]
    saved = []
    for r in resident:
        q, k, v = (r[x].detach() for x in ("query", "key", "value"))
        out_pre, den, scale, pos = ops.attention_fwd(d, q, k, v, r["coords"], w_rpe.weight.detach(), K, mod.e2lsh.alpha,
                                                     combined_shifts=r["combined_shifts"])
        saved.append((q, k, v, r["coords"], scale, pos, out_pre, den, torch.randn_like(out_pre)))

    def ev_time(fn, reps=8):
        for i in range(3):
            fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(reps):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps * 1e-3

    variant = lib.hept_get_bwd_variant()
    tc = bool(lib.hept_get_engine()) and variant >= 3
    t_fwd = ev_time(lambda i: ops.block_attention_fwd(d, *saved[i % n_sets][:6]))
    lib.hept_set_bwd_stage_mask(0)            # the streaming pre-pass alone (gradient rows, scaled coordinates)
    t_pre = ev_time(lambda i: ops.attention_bwd(d, *saved[i % n_sets]))
    lib.hept_set_bwd_stage_mask(3)            # pre-pass + tile kernel(s), without the final reductions
    t_tiles = ev_time(lambda i: ops.attention_bwd(d, *saved[i % n_sets]))
    lib.hept_set_bwd_stage_mask(7)
    t_bwd = max(t_tiles - t_pre, 1e-9)
    kernel = "block_attn_bwd_tc_kernel" if tc else "block_attn_bwd_dq_kernel + block_attn_bwd_dkv_kernel"
    bytes_per_launch = BWD_BYTES_PER_HIT * N_RAW
    achieved = bytes_per_launch / t_bwd / 1e9
    traffic, tsrc = None, None
    tj, tpath = _committed(PROFILE_TAG + "_traffic.json")
    if tj is None:
        tj, tpath = _committed("r1e_traffic.json")
    if tj and tc and "block_attn_bwd_tc_kernel" in tj:
        traffic = tj["block_attn_bwd_tc_kernel"]["dram_bytes_per_launch"]
        tsrc = f"{tpath} ({tj['block_attn_bwd_tc_kernel']['source']})"
    # tensor side.  No TF32 GEMM peak is in MEASURED_PEAKS.json: half the measured sustained bf16 rate is used and said so.
    peaks, _ = None, None
    ppath = os.path.join(ROOT, "MEASURED_PEAKS.json")
    bf16 = 1400.0
    if os.path.exists(ppath):
        with open(ppath) as f:
            bf16 = float(json.load(f).get("bf16_tflops_sustained", bf16))
    tf32_peak = bf16 / 2
    bwd_flop_per_hit = 2 * T * H * B * (2 * (D + C) + 2 * D + 2 * (D + C) + D)       # S, dP on both sides, dV, dQ, dK
    alg_tflops = 2 * T * H * B * (2 * (D + C) + 3 * D) * N_RAW / t_bwd / 1e12          # without the second evaluation of S, dP
    issued_tflops = 3 * bwd_flop_per_hit * N_RAW / t_bwd / 1e12                        # 3xTF32: three tensor-core passes each
    pad = (B * B) / (128.0 * ((B + 15) // 16 * 16))                                   # 100 x 100 tiles run as M = 128, N = 112
    ceil_hits = tf32_peak * 1e12 / 3 * pad / FLOP_PER_HIT
    return {
        "bound": "tensor" if tc else "fp32",
        "bound_evidence": ("ncu --set full of this kernel (profiles/): DRAM throughput ~15 % of peak, tensor pipe ~35 % active, issue slots "
                           "~39 %: no unit is saturated; the tensor pipe (3xTF32, M = 128 padding, 52-cycle floor of the N = 32 / 64 "
                           "products) is the nearest ceiling, the thread-side TMEM port what keeps it from it (DESIGN.md 4)"),
        "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "bytes_per_launch": bytes_per_launch, "bytes_per_hit": BWD_BYTES_PER_HIT,
        "bytes_model": "SURVEY.md 8(d) backward: dout, q, k, v, coords, permutations, saved (O, L), dq, dk, dv",
        "traffic": traffic, "traffic_source": tsrc, "peak_source": peak_src,
        "kernel_ms": {"block_attn_fwd": t_fwd * 1e3, "block_attn_bwd": t_bwd * 1e3, "bwd_pre_pass": t_pre * 1e3},
        "fwd_kernel": {"bytes_per_hit": FWD_BYTES_PER_HIT, "achieved": FWD_BYTES_PER_HIT * N_RAW / t_fwd / 1e9,
                       "frac": FWD_BYTES_PER_HIT * N_RAW / t_fwd / 1e9 / peak},
        "tensor_frac": issued_tflops / tf32_peak, "issued_tf32_tflops": issued_tflops, "algorithmic_tflops": alg_tflops,
        "tf32_peak_tflops": tf32_peak,
        "tf32_peak_source": "half of bf16_tflops_sustained in MEASURED_PEAKS.json (no TF32 GEMM was measured)",
        "ceiling": {"hits_per_s": ceil_hits, "frac_of_hbm_roofline": FWD_BWD_BYTES_PER_HIT * ceil_hits / 1e9 / peak,
                    "achieved_frac_of_ceiling": path_hits_per_s / ceil_hits,
                    "why": ("952 128 algorithmic FLOP per hit, each three TF32 tensor-core passes (3xTF32 keeps fp32-grade scores), "
                            "100 x 100 tiles issued as M = 128 x N = 112: the whole path cannot exceed this fraction of the HBM "
                            "roofline whatever the kernels do; the 60 % target assumes one pass per FLOP")},
    }


_JSON_FD = None


def emit(line: dict) -> None:
    """The ONE JSON line goes to the process's original stdout; everything else any library prints (NCCL's version banner,
    warnings) was routed to stderr by main()."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)          # keep the real stdout for the JSON line ...
    os.dup2(2, 1)                 # ... and send every other write to fd 1 (C libraries included) to stderr
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nccl", action="store_true", help="all-reduce the gradient bucket with NCCL instead of the library's NVLink kernel")
    ap.add_argument("--e2e-diag", action="store_true", help="extra end-to-end legs without H2D / D2H (diagnostic)")
    ap.add_argument("--no-train", action="store_true", help="skip the training-step leg (BASELINE.json configs[4])")
    ap.add_argument("--bwd", default=None, type=int, choices=[1, 3, 4, 5], help="backward tile variant (default: library default)")
    ap.add_argument("--engine", default=None, choices=["simt", "tcgen05"], help="tile engine (default: library default)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

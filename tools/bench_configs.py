"""Throughput of the other BASELINE.json configs (informational; bench.py carries the headline config).

    python tools/bench_configs.py                      # 1 GPU
    python -m torch.distributed.run --nproc-per-node N ... tools/bench_configs.py   # config 5 sharded by event

configs[1]  HEPTAttention fwd+bwd, 60 000 and 61 237 (ragged -> 61 300 padded) hits
configs[2]  pileup Transformer (src flavour, 4 layers), forward-only inference, 10 000-hit events
configs[3]  batched imbalanced events (8 events, 60 187 hits) through prepare_input + HEPTAttention fwd+bwd
configs[4]  tracking Transformer training step (fwd + bwd + flat NCCL all-reduce + Adam), one 60k event per rank
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from hept_b200 import HEPTAttention, prepare, sharding, synthetic
from hept_b200.model import Transformer

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
TRACKING = {k: v for k, v in synthetic.TRACKING.items() if k != "coords_dim"}
PILEUP = {k: v for k, v in synthetic.PILEUP.items() if k != "coords_dim"}


def timeit(fn, steps=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    return ms


def attention_problem(sizes, seed):
    cfg = dict(synthetic.TRACKING)
    coords, batch = synthetic.batched_cloud(sizes, 6, seed)
    params = synthetic.module_params(cfg, 0)
    helper = {"block_size": 100, "regions": params["regions"].to(dev), "num_heads": 8}
    _, kw, _ = prepare.prepare_input(torch.zeros(coords.shape[0], 1, device=dev), coords.to(dev), batch.to(dev), helper)
    n = kw["coords"].shape[0]
    q, k, v = (t.to(dev).requires_grad_(True) for t in synthetic.qkv(n, cfg, seed))
    mod = HEPTAttention(30, **cfg).to(dev)
    w_rpe = torch.nn.Linear(50, 192).to(dev)
    g = torch.randn(n, 24, device=dev)

    def step():
        for t in (q, k, v):
            t.grad = None
        mod(q, k, v, w_rpe=w_rpe, **kw).backward(g)

    def graphed():
        from hept_b200.graphed import graphed_attention

        gs = graphed_attention(mod, w_rpe, q, k, v, kw["coords"], kw["combined_shifts32"])

        def gstep():
            for t in (q, k, v):
                t.grad = None
            gs(q, k, v, kw["coords"], kw["combined_shifts32"]).backward(g)

        return gstep

    return step, n, graphed


results = []
if world == 1:
    for name, sizes in (("attention fwd+bwd, 6037 hits (tracking-6k)", [6037]), ("attention fwd+bwd, 60000 hits", [60000]), ("attention fwd+bwd, 61237 hits (padded to 61300)", [61237]),
                        ("attention fwd+bwd, 8 imbalanced events, 60187 hits", synthetic.event_sizes("batched-imbalanced"))):
        step, n, graphed = attention_problem(sizes, 3)
        ms = timeit(step)
        ms_g = timeit(graphed())
        results.append({"config": name, "ms_per_step": ms, "hits_per_s": sum(sizes) / ms * 1e3, "padded_hits": n,
                        "ms_per_step_cuda_graph": ms_g, "hits_per_s_cuda_graph": sum(sizes) / ms_g * 1e3})
    m = Transformer(in_dim=8, coords_dim=4, task="pileup", flavour="src", **PILEUP).eval().to(dev)
    for n in (5000, 10000, 20000):
        coords = synthetic.point_cloud(n, 4, 8).to(dev)
        x = torch.cat([torch.randn(n, 7) * 0.5, torch.randint(0, 7, (n, 1)).float()], dim=1).to(dev)
        with torch.no_grad():
            ms = timeit(lambda: m(x, coords))
        from hept_b200.graphed import GraphedInference

        gi = GraphedInference(m, (x, coords))
        ms_g = timeit(lambda: gi.run(x, coords))
        results.append({"config": f"pileup Transformer forward-only, {n} hits", "ms_per_step": ms, "hits_per_s": n / ms * 1e3,
                        "ms_per_step_cuda_graph": ms_g, "hits_per_s_cuda_graph": n / ms_g * 1e3})

# configs[4]: training step of the tracking model, one 60k event per rank per step
model = Transformer(in_dim=15, coords_dim=6, **TRACKING).to(dev)
opt = torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=1e-3)
coords = synthetic.point_cloud(60000, 6, 100 + rank).to(dev)
x = (torch.randn(60000, 15, generator=torch.Generator().manual_seed(rank)) * 0.5).to(dev)
grad_bytes = [0]


def train_step():
    opt.zero_grad(set_to_none=True)
    out = model(x, coords)
    loss = (out ** 2).mean()                       # stand-in loss: InfoNCE needs torch_scatter / pair lists (out of scope)
    loss.backward()
    grad_bytes[0] = sharding.allreduce_gradients(model.parameters())
    opt.step()


ms = timeit(train_step, steps=5, warmup=2)
results.append({"config": f"tracking Transformer training step, 60000 hits/rank, world={world}", "ms_per_step": ms,
                "hits_per_s": world * 60000 / ms * 1e3, "allreduce_bytes": grad_bytes[0]})
if rank == 0:
    print(json.dumps(results, indent=1))
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/configs_{world}gpu.json", "w") as f:
        json.dump(results, f, indent=1)
if world > 1:
    dist.destroy_process_group()

"""Under torchrun on >= 2 GPUs of one box: GradBucket(p2p=True) — the library's one-shot NVLink all-reduce — against NCCL on
the same data, and the time of both (CUDA events, max over ranks).

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/test_p2p.py
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from hept_b200 import sharding

rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
res = {}
for n in (14_352, 326_436, 5):                      # the attention module's bucket, the tracking model's, a tiny odd one
    params = [torch.nn.Parameter(torch.zeros(n - n // 3, device=dev)), torch.nn.Parameter(torch.zeros(n // 3, device=dev))]
    bucket = sharding.GradBucket(params, p2p=True)
    ref = torch.zeros(n, device=dev)
    ok = True
    for step in range(5):
        g = torch.Generator(device=dev).manual_seed(1000 * step + rank)
        vals = torch.randn(n, generator=g, device=dev)
        bucket.zero()
        bucket.flat += vals
        ref.copy_(vals)
        bucket.allreduce(average=True)
        dist.all_reduce(ref, op=dist.ReduceOp.AVG)
        ok &= bool(torch.allclose(bucket.flat, ref, rtol=1e-6, atol=1e-6))
        gathered = [torch.empty_like(bucket.flat) for _ in range(world)]
        dist.all_gather(gathered, bucket.flat.clone())
        ok &= all(torch.equal(gathered[0], t) for t in gathered)       # the same bits on every rank
    ok &= bucket.p2p_ok() and bucket.attached()

    def timeit(fn, reps=50):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps * 1e3], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    res[n] = {"p2p": bucket.uses_p2p, "p2p_error": bucket.p2p_error, "match": ok,
              "us_p2p": timeit(lambda: bucket.allreduce()), "us_nccl": timeit(lambda: dist.all_reduce(ref, op=dist.ReduceOp.AVG))}
if rank == 0:
    print(json.dumps(res))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open(f"gpurun_out/p2p_allreduce_{world}gpu.json", "w"), indent=1)
dist.barrier()
dist.destroy_process_group()
assert all(v["match"] for v in res.values()), res

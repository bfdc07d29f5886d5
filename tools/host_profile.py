"""Host-side cost of one small fwd+bwd (tracking-6k): cProfile of 300 steps, where the Python time goes.

    python tools/host_profile.py [n_hits]
"""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from hept_b200 import HEPTAttention, prepare, synthetic

n_hits = int(sys.argv[1]) if len(sys.argv) > 1 else 6037
dev = torch.device("cuda:0")
cfg = dict(synthetic.TRACKING)
coords, batch = synthetic.batched_cloud([n_hits], 6, 3)
params = synthetic.module_params(cfg, 0)
helper = {"block_size": 100, "regions": params["regions"].to(dev), "num_heads": 8}
_, kw, _ = prepare.prepare_input(torch.zeros(coords.shape[0], 1, device=dev), coords.to(dev), batch.to(dev), helper)
n = kw["coords"].shape[0]
q, k, v = (t.to(dev).requires_grad_(True) for t in synthetic.qkv(n, cfg, 3))
mod = HEPTAttention(30, **cfg).to(dev)
w_rpe = torch.nn.Linear(50, 192).to(dev)
g = torch.randn(n, 24, device=dev)


def step():
    for t in (q, k, v):
        t.grad = None
    mod(q, k, v, w_rpe=w_rpe, **kw).backward(g)


for _ in range(20):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(300):
    step()
t_issue = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
print(f"n={n}: host issue {t_issue / 300 * 1e6:.0f} us/step, wall {t_all / 300 * 1e6:.0f} us/step")
pr = cProfile.Profile()
pr.enable()
for _ in range(300):
    step()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)

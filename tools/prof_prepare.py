"""prepare_input on the library's kernels: wall time per call (one 60 000-hit event, the 8 imbalanced events of
BASELINE.json configs[3], one src/-flavour event) — the command ncu wraps for the prepare kernels."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from hept_b200 import prepare, synthetic

dev = torch.device("cuda:0")
params = synthetic.module_params(dict(synthetic.TRACKING), 0)
helper = {"block_size": 100, "regions": params["regions"].to(dev), "num_heads": 8}
for sizes in ([60000], synthetic.event_sizes("batched-imbalanced")):
    coords, batch = synthetic.batched_cloud(sizes, 6, 1)
    coords, batch = coords.to(dev), batch.to(dev)
    x = torch.zeros(coords.shape[0], 1, device=dev)
    for i in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        prepare.prepare_input(x, coords, batch, helper, sizes=sizes)
        torch.cuda.synchronize()
        print(len(sizes), "events: prepare_input wall ms", (time.perf_counter() - t0) * 1e3)
coords = synthetic.point_cloud(61237, 6, 2).to(dev)
x = torch.zeros(61237, 1, device=dev)
for i in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    prepare.prepare_input_single(x, coords, helper)
    torch.cuda.synchronize()
    print("single event: prepare_input_single wall ms", (time.perf_counter() - t0) * 1e3)

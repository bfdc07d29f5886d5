import sys, torch, time
sys.path.insert(0, "/root/repo")
from hept_b200 import synthetic, prepare
dev = torch.device("cuda:0")
params = synthetic.module_params(dict(synthetic.TRACKING), 0)
helper = {"block_size": 100, "regions": params["regions"].to(dev), "num_heads": 8}
for sizes in ([60000], synthetic.event_sizes("batched-imbalanced")):
    coords, batch = synthetic.batched_cloud(sizes, 6, 1)
    coords, batch = coords.to(dev), batch.to(dev); x = torch.zeros(coords.shape[0], 1, device=dev)
    for i in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter(); prepare.prepare_input(x, coords, batch, helper); torch.cuda.synchronize()
        print(len(sizes), "events: prepare_input wall ms", (time.perf_counter() - t0) * 1e3)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as p:
    prepare.prepare_input(x, coords, batch, helper); torch.cuda.synchronize()
print(p.key_averages().table(sort_by="cuda_time_total", row_limit=12, max_name_column_width=50))

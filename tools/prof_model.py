import os, sys, torch
sys.path.insert(0, "/root/repo")
from hept_b200 import synthetic, prepare
from hept_b200.model import Transformer
T = {k: v for k, v in synthetic.TRACKING.items() if k != "coords_dim"}
dev = torch.device("cuda:0")
model = Transformer(in_dim=15, coords_dim=6, **T).to(dev)
coords = synthetic.point_cloud(60000, 6, 1).to(dev); x = torch.randn(60000, 15, device=dev) * 0.5
def ev(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); [fn() for _ in range(n)]; e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
helper = dict(model.helper_params, regions=model.regions)
batch = torch.zeros(60000, dtype=torch.long, device=dev)
print("prepare_input ms", ev(lambda: prepare.prepare_input(x, coords, batch, helper)))
model.eval()
with torch.no_grad(): print("model fwd eval ms", ev(lambda: model(x, coords)))
model.train()
def step():
    model.zero_grad(set_to_none=True); (model(x, coords) ** 2).mean().backward()
print("model fwd+bwd ms", ev(step))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as p:
    step(); torch.cuda.synchronize()
print(p.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=60))

import sys; sys.path.insert(0, "/root/repo")
import torch
from hept_b200 import synthetic, metrics
dev = "cuda:0"
n = 60000
cid, recons, pts, pairs = (a.to(dev) for a in synthetic.tracking_truth(n, 5))
emb = (torch.randn(n, 12, generator=torch.Generator().manual_seed(9)) * 0.5).to(dev)
mask = metrics.point_filter(cid, recons, pts, 0.9)
for i in range(2):
    print(metrics.acc_and_pr_at_k(emb, cid, mask, "l2_rbf", K=31))

"""Golden fixture for the InfoNCE loss: the reference's ``InfoNCELoss`` (src/utils/losses.py) run in this
container on seeded inputs.

    python tests/golden/make_loss_fixture.py      # writes tests/golden/infonce_small.npz

The reference's loss imports ``torch_scatter`` (``segment_csr``, ``scatter_mean``), which is not installed here and cannot be
(no network).  Those two third-party functions are supplied by a stand-in module with their documented semantics
(segment_csr(src, indptr, reduce) = reduce over src[indptr[g]:indptr[g+1]]); every line of the reference's own files runs
unmodified.  The fixture therefore pins the oracle to "the reference + a restated third-party segment reduction".
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("HEPT_REFERENCE", "/root/reference")


def load_reference_losses():
    ts = types.ModuleType("torch_scatter")

    def segment_csr(src, indptr, out=None, reduce="sum"):
        res = []
        for g in range(indptr.numel() - 1):
            seg = src[int(indptr[g]): int(indptr[g + 1])]
            res.append(seg.mean() if reduce == "mean" else seg.sum())
        return torch.stack(res) if res else src.new_zeros(0)

    def scatter_mean(src, index, dim=-1, out=None, dim_size=None):
        n = int(index.max()) + 1 if dim_size is None else dim_size
        s = torch.zeros(n, dtype=src.dtype).index_add_(0, index, src)
        c = torch.zeros(n, dtype=src.dtype).index_add_(0, index, torch.ones_like(src))
        return s / c.clamp(min=1)

    ts.segment_csr, ts.scatter_mean = segment_csr, scatter_mean
    sys.modules["torch_scatter"] = ts
    pkg = types.ModuleType("refutils")
    pkg.__path__ = [os.path.join(REF, "src", "utils")]
    sys.modules["refutils"] = pkg
    mods = {}
    for name in ("metrics", "losses"):
        spec = importlib.util.spec_from_file_location(f"refutils.{name}", os.path.join(REF, "src", "utils", f"{name}.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[f"refutils.{name}"] = mod
        spec.loader.exec_module(mod)
        mods[name] = mod
    return mods["losses"], mods["metrics"]


def problem(n=1500, dim=12, seed=0, per_point=24):
    """Seeded embeddings, cluster ids (64-bit, some 0), reconstructable flags, pt values and point pairs: for every point its
    cluster mates plus ``per_point`` random others (so every point owns negative pairs), shuffled."""
    g = torch.Generator().manual_seed(seed)
    n_part = n // 9
    owner = torch.randint(0, n_part, (n,), generator=g)
    centre = torch.randn(n_part, dim, generator=g)
    x = (centre[owner] + 0.15 * torch.randn(n, dim, generator=g)).float().contiguous()
    cid = (owner.long() + 1) * 4503599627370497 % (1 << 62)
    cid[torch.rand(n, generator=g) < 0.05] = 0
    recons = (torch.rand(n, generator=g) < 0.9).float()
    pts = torch.rand(n, generator=g) * 3.0
    src = torch.arange(n).repeat_interleave(per_point)
    dst = torch.randint(0, n, (n * per_point,), generator=g)
    mates = (owner[:, None] == owner[None, :]).nonzero().T
    pairs = torch.cat([torch.stack([src, dst]), mates], dim=1)
    pairs = pairs[:, pairs[0] != pairs[1]]
    pairs = pairs[:, torch.randperm(pairs.shape[1], generator=g)].contiguous()
    return x, pairs, cid, recons, pts


def metric_problem(n=1800, dim=12, seed=1):
    """Embeddings clustered by particle (clusters of 2 .. 14 points, some singletons), cluster ids, and a point mask."""
    g = torch.Generator().manual_seed(seed)
    sizes = torch.randint(1, 15, (n,), generator=g)
    owner = torch.repeat_interleave(torch.arange(n), sizes)[:n]
    owner = owner[torch.randperm(n, generator=g)]
    centre = torch.randn(int(owner.max()) + 1, dim, generator=g) * 2.0
    x = (centre[owner] + 0.3 * torch.randn(n, dim, generator=g)).float().contiguous()
    cid = (owner.long() + 1) * 4503599627370497 % (1 << 62)
    mask = torch.rand(n, generator=g) < 0.6
    return x, cid, mask


def main():
    L, M = load_reference_losses()
    out = {}
    x, cid, mask = metric_problem()
    for metric in ("l2_rbf", "cosine"):
        out[f"knn_{metric}"] = np.asarray(M.acc_and_pr_at_k(x, cid, mask, metric, K=19), dtype=np.float64)
    out["meta_chk_knn_x"] = np.asarray(float(x.double().sum()))
    for metric in ("l2_rbf", "l2_inverse", "cosine"):
        x, pairs, cid, recons, pts = problem()
        xr = x.clone().requires_grad_(True)
        loss = L.InfoNCELoss(tau=0.05, dist_metric=metric)(xr, pairs, cid, recons, pts)
        loss.backward()
        out[f"loss_{metric}"] = loss.detach().numpy()
        out[f"dx_{metric}"] = xr.grad.numpy()
    out["meta_chk_x"] = np.asarray(float(x.double().sum()))
    out["meta_num_pairs"] = np.asarray(pairs.shape[1])
    path = os.path.join(HERE, "infonce_small.npz")
    np.savez_compressed(path, **out)
    print({k: (float(v) if v.ndim == 0 else v.shape) for k, v in out.items()}, os.path.getsize(path))


if __name__ == "__main__":
    main()

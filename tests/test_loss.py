"""the InfoNCE loss of the tracking task.

CPU: the oracle's restatement (oracle/hept_oracle.py::infonce_loss) against the golden fixture produced by the reference's
own InfoNCELoss (tests/golden/make_loss_fixture.py; torch_scatter supplied by a stand-in).  GPU: the library's kernels
(hept_infonce_fwd / hept_infonce_bwd through the C ABI, hept_b200.losses.InfoNCELoss) against the oracle in float64:
    |loss - loss64| <= 2.5 |loss32 - loss64| + 2e-6 |loss64|,   rel. Frobenius error of d x likewise with floor 1e-5.
"""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import hept_oracle as O
from tests.helpers import GOLDEN, rel_err

sys.path.insert(0, GOLDEN)
import make_loss_fixture as MLF  # noqa: E402  (only its seeded problem generator; the reference is not imported by this)

METRICS = ("l2_rbf", "l2_inverse", "cosine")


def _oracle(x, pairs, cid, recons, pts, metric, dtype, compact=True):
    xr = x.detach().to(dtype).clone().requires_grad_(True)
    loss = O.infonce_loss(xr, pairs, cid, recons.to(dtype), pts.to(dtype), 0.05, metric, compact_like_reference=compact)
    loss.backward()
    return loss.detach(), xr.grad


@pytest.mark.parametrize("metric", METRICS)
def test_oracle_matches_reference_fixture(metric):
    z = np.load(os.path.join(GOLDEN, "infonce_small.npz"))
    x, pairs, cid, recons, pts = MLF.problem()
    assert abs(float(x.double().sum()) - float(z["meta_chk_x"])) < 1e-9 * abs(float(z["meta_chk_x"]))
    loss, dx = _oracle(x, pairs, cid, recons, pts, metric, torch.float32)
    assert abs(float(loss) - float(z[f"loss_{metric}"])) <= 2e-6 * abs(float(z[f"loss_{metric}"]))
    assert rel_err(dx, torch.from_numpy(z[f"dx_{metric}"])) < 1e-5
    # every point owns a negative pair here, so indexing by point number (the product's convention) changes nothing
    loss2, dx2 = _oracle(x, pairs, cid, recons, pts, metric, torch.float32, compact=False)
    assert torch.equal(loss, loss2) and rel_err(dx2, dx) < 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("metric", METRICS)
@pytest.mark.parametrize("n,per_point", [(1500, 24), (20000, 40), (300, 3)])
def test_infonce_kernels_against_oracle(metric, n, per_point):
    from hept_b200.losses import InfoNCELoss

    x, pairs, cid, recons, pts = MLF.problem(n=n, seed=n, per_point=per_point)
    l32, d32 = _oracle(x, pairs, cid, recons, pts, metric, torch.float32, compact=False)
    l64, d64 = _oracle(x, pairs, cid, recons, pts, metric, torch.float64, compact=False)
    dev = "cuda:0"
    crit = InfoNCELoss(tau=0.05, dist_metric=metric)
    xd = x.to(dev).requires_grad_(True)
    args = (pairs.to(dev), cid.to(dev), recons.to(dev), pts.to(dev))
    loss = crit(xd, *args)
    loss.backward()
    assert abs(float(loss) - float(l64)) <= 2.5 * abs(float(l32) - float(l64)) + 2e-6 * abs(float(l64))
    e_o, e_r = rel_err(xd.grad.cpu(), d64), rel_err(d32, d64)
    assert e_o <= 2.5 * e_r + 1e-5, (e_o, e_r)
    # deterministic: the CSR fill order comes from atomics, the summation order does not
    x2 = x.to(dev).requires_grad_(True)
    loss2 = crit(x2, *args)
    loss2.backward()
    assert torch.equal(loss, loss2) and torch.equal(xd.grad, x2.grad)


@pytest.mark.gpu
def test_infonce_against_reference_fixture_on_gpu():
    from hept_b200.losses import InfoNCELoss

    z = np.load(os.path.join(GOLDEN, "infonce_small.npz"))
    x, pairs, cid, recons, pts = MLF.problem()
    for metric in METRICS:
        xd = x.to("cuda:0").requires_grad_(True)
        loss = InfoNCELoss(0.05, metric)(xd, pairs.cuda(), cid.cuda(), recons.cuda(), pts.cuda())
        loss.backward()
        assert abs(float(loss) - float(z[f"loss_{metric}"])) <= 1e-5 * abs(float(z[f"loss_{metric}"]))
        assert rel_err(xd.grad.cpu(), torch.from_numpy(z[f"dx_{metric}"])) < 1e-4


# ------------------------------------------------------------------------------------------ kNN metrics
@pytest.mark.parametrize("metric", ["l2_rbf", "cosine"])
def test_oracle_knn_metrics_match_reference_fixture(metric):
    z = np.load(os.path.join(GOLDEN, "infonce_small.npz"))
    x, cid, mask = MLF.metric_problem()
    assert abs(float(x.double().sum()) - float(z["meta_chk_knn_x"])) < 1e-9 * abs(float(z["meta_chk_knn_x"]))
    got = O.knn_metrics(x, cid, mask, metric, K=19)
    assert np.allclose(np.asarray(got), z[f"knn_{metric}"], rtol=0, atol=1e-9)


@pytest.mark.gpu
@pytest.mark.parametrize("metric", ["l2_rbf", "cosine"])
@pytest.mark.parametrize("n", [1800, 20000])
def test_knn_metrics_kernel_against_oracle(metric, n):
    """hept_knn_metrics against the oracle (dense cdist + topk): the neighbour SETS can differ only where two candidates are
    equally far to rounding (cdist takes the matmul form of the distance; 1 - cos carries ~1e-7 of noise), so the three means
    agree to a couple of queries' worth of neighbours."""
    from hept_b200 import metrics

    x, cid, mask = MLF.metric_problem(n=n, seed=n)
    want = O.knn_metrics(x, cid, mask, metric, K=19)
    got = metrics.acc_and_pr_at_k(x.cuda(), cid.cuda(), mask.cuda(), metric, K=19)
    assert np.allclose(np.asarray(got), np.asarray(want), rtol=0, atol=max(1e-4, 2.0 / int(mask.sum()))), (got, want)
    again = metrics.acc_and_pr_at_k(x.cuda(), cid.cuda(), mask.cuda(), metric, K=19)
    assert got == again


@pytest.mark.gpu
def test_knn_metrics_against_reference_fixture_on_gpu():
    from hept_b200 import metrics

    z = np.load(os.path.join(GOLDEN, "infonce_small.npz"))
    x, cid, mask = MLF.metric_problem()
    for metric in ("l2_rbf", "cosine"):
        got = metrics.acc_and_pr_at_k(x.cuda(), cid.cuda(), mask.cuda(), metric, K=19)
        assert np.allclose(np.asarray(got), z[f"knn_{metric}"], rtol=0, atol=max(1e-4, 2.0 / int(mask.sum())))

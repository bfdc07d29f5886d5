"""The tracking task's kNN metrics on the library's kernels: ``acc_and_pr_at_k`` / ``point_filter`` of the
reference (src/utils/metrics.py:19-62, scoring loop :65-93) with the same signatures.

The dense (queries x N) distance matrix, the top-k and the numba scoring loop of the reference become one streaming kernel
(``hept_knn_metrics``, csrc/metrics.cu); only the three means come back to the host.  Ties between equal distances are broken
by the lower point index (torch.topk does not specify an order).  There is no CPU path.
"""
from __future__ import annotations

import torch

from . import ops


def point_filter(cluster_ids, recons, pts, pt_thres):
    return (cluster_ids != 0) & (recons != 0) & (pts > pt_thres)


@torch.no_grad()
def acc_and_pr_at_k(embeddings, cluster_ids, mask, dist_metric, K=19, batch_size=None):
    """-> (accuracy, precision, recall) as Python floats.  ``batch_size`` is accepted for signature compatibility: the kernel
    never materialises the distance matrix, so there is nothing to batch."""
    if not embeddings.is_cuda:
        raise RuntimeError("hept_b200.metrics.acc_and_pr_at_k runs on CUDA (sm_100a) only; there is no CPU path")
    if "l2" in dist_metric:
        cosine = False
    elif dist_metric == "cosine":
        cosine = True
    else:
        raise NotImplementedError(dist_metric)
    queries = torch.nonzero(mask.to(embeddings.device), as_tuple=False).flatten()
    out = ops.knn_metrics(embeddings.float().contiguous(), cluster_ids.long().contiguous(), queries, cosine, K).tolist()
    assert out[4] <= K, f"K is too small, max k is {int(out[4])}"
    return out[0], out[1], out[2]

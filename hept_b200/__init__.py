"""hept_b200 — B200-native (sm_100a) implementation of HEPT's LSH-bucketed attention.

Public surface (mirrors the reference, Graph-COM/HEPT):
  HEPTAttention        drop-in for example/hept.py:31-81 and src/models/attention/hept.py:59-117
  prepare              per-forward hash-code preparation (example/transformer.py:35-63 and the HEPT
                       branch of src/models/baselines/transformer.py:43-57)
  attention.attn_front the front of the Attn block, norm1 -> w_q / w_k / w_v (example/transformer.py:157-158)
  losses.InfoNCELoss, metrics.acc_and_pr_at_k    the tracking task's loss and kNN metrics (src/utils/losses.py, metrics.py)
  graphed              CUDA-graph replay of the module call for fixed shapes
  ops                  stage-wise wrappers over the C ABI in include/hept_b200.h
"""
from .attention import HEPTAttention, E2LSH  # noqa: F401
from . import ops  # noqa: F401

__all__ = ["HEPTAttention", "E2LSH", "ops"]

"""CUDA-event time of the segmented argsort alone (a7), tracking-60k shape by default; inputs rotate over 4 key sets.

    [HEPT_SORT_CLUSTER=cs] python tools/sort_times.py [segments] [n]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from hept_b200 import _lib, ops

segs = int(sys.argv[1]) if len(sys.argv) > 1 else 48
n = int(sys.argv[2]) if len(sys.argv) > 2 else 60000
lib = _lib.load()
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(1)
keys = [(torch.randn(2, segs // 2, n, generator=g) * 50 + torch.randint(0, 40, (2, segs // 2, n), generator=g) * 400.0).to(dev)
        for _ in range(4)]


def ev(reps=20):
    for i in range(3):
        ops.segmented_argsort(keys[i % 4])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        ops.segmented_argsort(keys[i % 4])
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for variant in (0, 1):
    lib.hept_set_sort_variant(variant)
    pos = ops.segmented_argsort(keys[0])
    ok = torch.equal(pos.long(), torch.argsort(keys[0], dim=-1, stable=True))
    print(f"variant {variant} cluster={os.environ.get('HEPT_SORT_CLUSTER', 'auto')}: {ev():.1f} us, stable argsort: {ok}")
lib.hept_set_sort_variant(0)

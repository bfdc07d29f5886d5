"""Cost model of small tcgen05.mma bursts on this GPU (cycles from first issue to commit completion).

    python tools/umma_timing.py      -> gpurun_out/umma_timing.json
"""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tests import native as _lib

lib = _lib.load()
out = torch.zeros(1, dtype=torch.int64, device="cuda:0")
res = {}
names = {0: "ts_n32_1acc", 1: "ts_n32_2acc", 5: "ts_n64", 6: "ts_n96", 7: "ts_n128", 8: "ss_n32", 2: "ss_n112_1acc",
         9: "bwd_dV_placement(A 224/336, D 448)", 10: "bwd_dK_placement(A 0/112, D 448)", 11: "bwd_dQ_placement(A 0/112, D 224)"}
for mode, name in names.items():
    for count in ((2, 26, 52) if mode >= 9 else (1, 13, 39)):
        _lib.check(lib.hept_debug_umma_timing(mode, count, ctypes.c_void_p(out.data_ptr()),
                                              ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "umma_timing")
        torch.cuda.synchronize()
        res[f"{name}_x{count}"] = int(out.item())
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/umma_timing.json", "w"), indent=1)
for k, v in res.items():
    print(k, v)

"""The library's one-shot NVLink all-reduce of the gradient bucket (csrc/p2p.cu, sharding.GradBucket(p2p=True)) against NCCL:
needs two GPUs on the box (skipped on the single-GPU box the suite usually runs on; `tools/test_p2p.py` is the same check
under torchrun and was run on 2 and 8 B200s: gpurun_out/p2p_allreduce_*gpu.json)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_p2p_allreduce_matches_nccl_on_two_gpus():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29541", os.path.join(root, "tools", "test_p2p.py")], cwd=root, capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]

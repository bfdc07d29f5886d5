"""Build libhept_sm100.so in-tree with nvcc for sm_100a (no JIT cache, no torch C++ headers).

    python -m hept_b200.build [--force] [--verbose]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libhept_sm100.so")
    if force:
        subprocess.run(["make", "-C", CSRC, "clean"], check=True, capture_output=not verbose)
    cmd = ["make", "-C", CSRC, f"-j{os.cpu_count() or 4}", f"NVCC={nvcc}"]
    if verbose:
        cmd.append("EXTRA=-Xptxas -v")
    res = subprocess.run(cmd, capture_output=not verbose, text=True)
    if res.returncode != 0:
        sys.stderr.write((res.stdout or "") + (res.stderr or ""))
        raise RuntimeError("building libhept_sm100.so failed")
    out = os.path.join(HERE, "libhept_sm100.so")
    assert os.path.exists(out), out
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))

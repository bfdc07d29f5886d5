#!/bin/bash
# compute-sanitizer passes over the GPU parity suite (run on a B200 box):
#   memcheck over everything, racecheck over the streaming kernels (the tcgen05 tile kernels synchronise through
#   mbarriers and TMEM, which racecheck does not model).  Round 1: 0 errors / 0 hazards.
set -e
cd "$(dirname "$0")/.."
compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py tests/test_gpu_umma.py -x -q
compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -x -q \
    -k "out_linear or projection or coord_scale or argsort"

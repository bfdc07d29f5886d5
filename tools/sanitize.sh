#!/bin/bash
# compute-sanitizer passes over a subset of the GPU suite small enough for the tool's slowdown (run on a B200 box):
#   memcheck over the small fixtures of the parity suite, the prepare kernels, the Attn front and the tcgen05 self-tests;
#   racecheck over the streaming kernels (the tcgen05 tile kernels synchronise through mbarriers and TMEM, which racecheck
#   does not model).  Logs go to gpurun_out/sanitizer_*.log; copies are committed under profiles/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py tests/test_gpu_umma.py tests/test_prepare.py \
    tests/test_attn_front.py tests/test_loss.py -x -q -m gpu \
    -k "(tiny_example or tiny_src or small_batched or small_src or umma or 57 or 130 or 127 or 1300 or 300-3 or 1800 or fixture_on_gpu) and not headline and not cluster_sizes" \
    > gpurun_out/sanitizer_memcheck.log 2>&1
echo "memcheck exit $?" >> gpurun_out/sanitizer_memcheck.log
compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py tests/test_prepare.py tests/test_attn_front.py tests/test_loss.py -x -q -m gpu \
    -k "(out_linear and 1300) or (projection and small_batched) or coord_scale or (argsort and 4097) or 130 or (attn_front and 1300) or 300-3 or (knn and 1800)" \
    > gpurun_out/sanitizer_racecheck.log 2>&1
echo "racecheck exit $?" >> gpurun_out/sanitizer_racecheck.log
tail -n 6 gpurun_out/sanitizer_memcheck.log gpurun_out/sanitizer_racecheck.log

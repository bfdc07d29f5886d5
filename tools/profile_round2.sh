#!/bin/bash
# All ncu captures of round 2 in one GPU call; summaries are made here afterwards with tools/ncu_summary.py.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
# (1) launch list of the bench command itself (device time of every launch; compare SHARES, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-train > gpurun_out/r2_bench_under_ncu.log 2>&1
# (2) launch list of one step at the Attn-block boundary through the stage-wise ABI (60k hits, then 6 037 hits)
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_launches_block.csv \
    python tools/profile_block.py 60000 2 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_launches_block_6k.csv \
    python tools/profile_block.py 6037 2 > /dev/null 2>&1
# (3) --set full of the two tile kernels (second repetition) and of the new streaming kernels
ncu --set full --clock-control none --import-source on -k regex:block_attn -s 2 -c 2 -f -o gpurun_out/prof_tiles_r2 \
    python tools/profile_step.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"rows_wide|rows_narrow|qkv_weights|params_mma|ln_params" -s 7 -c 8 -f -o gpurun_out/prof_front_r2 \
    python tools/profile_block.py 60000 2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"prep_|cluster_sort" -c 12 -f -o gpurun_out/prof_prepare_r2 \
    python tools/prof_prepare.py > /dev/null 2>&1
python tools/stage_times.py
ls -la gpurun_out/*.ncu-rep gpurun_out/r2_*

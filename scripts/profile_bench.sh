#!/bin/bash
# Usage (on the GPU box, under gpurun): bash scripts/profile_bench.sh <tag> [extra bench args]
# 1) launch list with per-kernel device time of one short bench run   2) ncu --set full of the hot kernels
set -x
TAG=${1:-r01}; shift
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline "$@" > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'update_tile_kernel|loss_thread_kernel|rollout_kernel|process_samples_kernel|lfb_gram' \
    -s 6 -c 6 -o gpurun_out/prof_${TAG} -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline "$@" > gpurun_out/ncu_full_${TAG}.log 2>&1
ls -la gpurun_out

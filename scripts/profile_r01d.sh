#!/bin/bash
# r01d: launch list of the default bench at the end-of-round state + ncu --set full of the 64-wide tiled-GEMM kernels
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01d.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu_r01d.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'update_gemm_kernel' -s 12 -c 2 -o gpurun_out/prof_r01d_hopper -f \
    python bench.py --workload hopper_trpo_4096x500 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_r01d.log 2>&1
ls -la gpurun_out | tail -8

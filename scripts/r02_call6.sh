#!/bin/bash
mkdir -p gpurun_out
{
for sd in 7 11 12 13 14; do timeout 300 python scripts/exp_fvp_precision.py f64 f64 f64 $sd 2>&1 | tail -1; done
for rep in 1 2; do for sd in 7 11 12; do echo "kernel f64 seed $sd rep $rep"; timeout 600 python scripts/swimmer_curve_gpu.py 40 swimmer f64 $sd 2>&1 | tail -1; done; done
for sd in 7 11 12; do echo "hopper torch-f64 seed $sd"; done
} | tee gpurun_out/r02_exp_seeds.log
echo "== pytest subset"; timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_algos.py -q -m gpu --tb=short -p no:cacheprovider -k "masking or f64_mode" 2>&1 | tail -15

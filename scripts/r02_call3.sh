#!/bin/bash
mkdir -p gpurun_out
echo "== umma modes"; timeout 180 python scripts/umma_modes_probe.py 2>&1 | tee gpurun_out/r02_umma_modes.log
for sd in 7 11 12; do
echo "== swimmer f32 round-p seed $sd"; EXP_ROUND_P=1 timeout 600 python scripts/swimmer_curve_gpu.py 40 swimmer f32 $sd 2>&1 | tee gpurun_out/r02_swimmer_rp_s$sd.log | tail -2
done
echo "== swimmer f64 seed 12"; timeout 900 python scripts/swimmer_curve_gpu.py 40 swimmer f64 12 2>&1 | tee gpurun_out/r02_swimmer_f64_s12.log | tail -2

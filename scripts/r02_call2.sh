#!/bin/bash
mkdir -p gpurun_out
echo "== umma modes"; timeout 180 python scripts/umma_modes_probe.py 2>&1 | tee gpurun_out/r02_umma_modes.log
echo "== swimmer f64"; timeout 900 python scripts/swimmer_curve_gpu.py 40 swimmer f64 2>&1 | tee gpurun_out/r02_swimmer_f64.log | tail -4
echo "== swimmer f32 seed 11"; timeout 600 python scripts/swimmer_curve_gpu.py 40 swimmer f32 11 2>&1 | tee gpurun_out/r02_swimmer_s11.log | tail -3
echo "== swimmer f32 seed 12 pseed 5"; timeout 600 python scripts/swimmer_curve_gpu.py 40 swimmer f32 12 5 2>&1 | tee gpurun_out/r02_swimmer_s12.log | tail -3
echo "== swimmer f64 seed 11"; timeout 900 python scripts/swimmer_curve_gpu.py 40 swimmer f64 11 2>&1 | tee gpurun_out/r02_swimmer_f64_s11.log | tail -3

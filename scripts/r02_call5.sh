#!/bin/bash
mkdir -p gpurun_out
for cfg in "f64 f64 f64" "f32 f32 f32" "f32 f64 f64" "f64 f32 f64" "f64 f64 f32"; do
  for sd in 7 11 12; do timeout 300 python scripts/exp_fvp_precision.py $cfg $sd 2>&1 | tail -1; done
done | tee gpurun_out/r02_exp_fvp_precision.log

#!/bin/bash
# which part of a float32 Fisher-vector product costs TRPO its learning speed on Swimmer? (torch FVP with selectable precision)
mkdir -p gpurun_out
for cfg in "f64 f64 f64" "f32 f32 f32" "f32 f64 f64" "f64 f32 f64" "f64 f64 f32" "f32 f32 f64"; do
  timeout 300 python scripts/exp_fvp_precision.py $cfg 7 2>&1 | tail -1
done | tee gpurun_out/r02_exp_fvp_precision.log

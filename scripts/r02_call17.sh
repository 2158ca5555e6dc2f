#!/bin/bash
# which float32 pass (gradient / Fisher-vector product / line-search loss) costs TRPO its learning speed on Swimmer?
mkdir -p gpurun_out
for sd in 7 11; do
for cfg in "f32 f32 f32" "f64 f32 f32" "f32 f64 f32" "f32 f32 f64" "f64 f64 f32" "f64 f64 f64"; do
  timeout 400 python scripts/exp_pass_precision.py $cfg $sd 2>&1 | tail -1
done; done | tee gpurun_out/r02_exp_pass_precision.log

#!/bin/bash
# Swimmer learning curves on the device (f32 and f64-CG) for the three sampler seeds of the committed oracle curves
mkdir -p gpurun_out
for prec in f32 f64; do for sd in 7 11 12; do
  echo "== $prec seed $sd"; timeout 600 python scripts/swimmer_curve_gpu.py 40 swimmer $prec $sd 2>&1 | tee gpurun_out/r02_swimmer_${prec}_s$sd.log | awk 'NR%4==1 || /mean Average/' | cut -c1-110
done; done

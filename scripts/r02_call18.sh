#!/bin/bash
mkdir -p gpurun_out
for sd in 7 11 12; do for v in base roundp hilo; do
  timeout 300 python scripts/exp_fvp_hilo.py $v $sd 2>&1 | tail -1
done; done | tee gpurun_out/r02_exp_fvp_hilo.log

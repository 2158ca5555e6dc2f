#!/bin/bash
# round-2 GPU call 1: tcgen05 probe, build-variant A/B (cw / ft / cwft), Swimmer + Hopper learning curves vs the oracle
mkdir -p gpurun_out
echo "== umma probe"; timeout 180 python scripts/umma_probe.py 2>&1 | tee gpurun_out/r02_umma_probe.log
echo "== swimmer curve"; timeout 600 python scripts/swimmer_curve_gpu.py 40 swimmer 2>&1 | tee gpurun_out/r02_swimmer_curve.log | tail -8
echo "== hopper curve"; timeout 600 python scripts/swimmer_curve_gpu.py 40 hopper 2>&1 | tee gpurun_out/r02_hopper_curve.log | tail -8
echo "== A/B"; timeout 1500 bash scripts/ab_const_weights.sh 2>&1 | tee gpurun_out/r02_ab_variants.log

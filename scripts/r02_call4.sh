#!/bin/bash
mkdir -p gpurun_out
echo "== umma fvp check"; timeout 300 python scripts/umma_fvp_check.py 2>&1 | tail -30 | tee gpurun_out/r02_umma_fvp.log
echo "== pytest gpu"; timeout 1700 python -m pytest tests -q -m gpu 2>&1 | tail -40 | tee gpurun_out/r02_pytest4.log
echo "== bench"; timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -3 | tee gpurun_out/r02_bench4.log

#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1700 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/r02_pytest4.log 2>&1; tail -25 gpurun_out/r02_pytest4.log
echo "== bench"; timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -3 | tee gpurun_out/r02_bench4.log
bash scripts/r02_call5.sh

#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fullsize.py tests/test_gpu_algos.py -q -m gpu --tb=short -p no:cacheprovider -x 2>&1 | tail -5
echo "== bench"; timeout 600 python bench.py --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], {k:v['ms'] for k,v in d['kernels'].items()}, {k:(v['ms_per_step']) for k,v in d['extra']['workloads'].items()})"

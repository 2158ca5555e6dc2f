#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fullsize.py -q -m gpu --tb=short -p no:cacheprovider -x 2>&1 | tail -3
for wl in swimmer_trpo_16384x500; do
timeout 300 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline --no-extra 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], {k:v['ms'] for k,v in d['kernels'].items()})"
done
timeout 300 python bench.py --workload point_trpo_65536x100 --steps 5 --warmup 3 --no-cpu-baseline --no-extra 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], {k:v['ms'] for k,v in d['kernels'].items()})"

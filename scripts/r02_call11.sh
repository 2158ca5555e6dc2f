#!/bin/bash
# 1 GPU: full GPU suite (incl. learning curves), bench, per-workload kernel timings
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 2400 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/r02_pytest11.log 2>&1; tail -25 gpurun_out/r02_pytest11.log
echo "== bench"; timeout 600 python bench.py 2> gpurun_out/r02_bench11.err | tail -1 > gpurun_out/r02_bench11.json; python -c "
import json
d=json.load(open('gpurun_out/r02_bench11.json')); print(d['ms_per_step'], d['value'], {k:v['ms'] for k,v in d['kernels'].items()}, d['roofline']['frac'], d['roofline_hbm']['frac'], d['extra'])"
tail -3 gpurun_out/r02_bench11.err
for wl in swimmer_trpo_16384x500 hopper_trpo_4096x500; do
  echo "== $wl"
  timeout 300 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu-baseline --no-extra 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], {k:v['ms'] for k,v in d['kernels'].items()}, d['stats'])"
done

#!/bin/bash
# last verification of the committed tree: full GPU suite, smoke, default bench
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 2400 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/r02_pytest26.log 2>&1; tail -6 gpurun_out/r02_pytest26.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== bench"; timeout 600 python bench.py 2> gpurun_out/r02_bench26.err | tail -1 > gpurun_out/r02_bench26.json; python -c "
import json
d=json.load(open('gpurun_out/r02_bench26.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches'], {k:v['ms'] for k,v in d['kernels'].items()}, d['roofline']['frac'], d['roofline_hbm']['frac'], d['clocks'], d['cpu_baseline']['value'], {k:(v['ms_per_step'],v['AverageReturn']) for k,v in d['extra']['workloads'].items()})"
tail -2 gpurun_out/r02_bench26.err

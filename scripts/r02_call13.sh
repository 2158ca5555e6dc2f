#!/bin/bash
# tensor-core loss pass, wide finalize, vectorised centering / baseline Gram: tests + bench
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 2400 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider -x > gpurun_out/r02_pytest13.log 2>&1; tail -25 gpurun_out/r02_pytest13.log
echo "== bench"; timeout 600 python bench.py 2> gpurun_out/r02_bench13.err | tail -1 > gpurun_out/r02_bench13.json; python -c "
import json
d=json.load(open('gpurun_out/r02_bench13.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches'], {k:v['ms'] for k,v in d['kernels'].items()}, d['roofline']['frac'], d['roofline_hbm']['frac'], {k:(v['ms_per_step'],v['AverageReturn']) for k,v in d['extra']['workloads'].items()})"
tail -3 gpurun_out/r02_bench13.err
echo "== launch list"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02c.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/bench_under_ncu_r02c.log 2>&1
python - <<'PY'
import csv,io
rows=[l for l in open('gpurun_out/launches_r02c.csv') if not l.startswith('==')]
r=list(csv.reader(io.StringIO(''.join(rows))))
h=r[0]; ni=h.index('Kernel Name'); vi=h.index('Metric Value')
agg={}
for x in r[1:]:
    try: v=float(x[vi].replace(',',''))
    except: continue
    a=agg.setdefault(x[ni][:70],[0,0.0]); a[0]+=1; a[1]+=v
for k,(n,v) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:16]: print('%-72s %4d %10.1f us  %.1f us/launch'%(k,n,v/1e3,v/1e3/n))
PY

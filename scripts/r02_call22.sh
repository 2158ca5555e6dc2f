#!/bin/bash
# float64-chain Fisher-vector product: parity, timing, Swimmer seed sweep with the new default
mkdir -p gpurun_out
echo "== pytest"; timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_round2.py tests/test_gpu_algos.py -q -m gpu --tb=short -p no:cacheprovider -x -k "not learning_curve" 2>&1 | tail -12
echo "== swimmer bench"; timeout 300 python bench.py --workload swimmer_trpo_16384x500 --steps 5 --warmup 3 --no-cpu-baseline --no-extra 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], {k:(v['ms'],v['frac']) for k,v in d['kernels'].items()})"
echo "== sweep (default = chain64)"; timeout 900 python scripts/exp_seed_sweep.py 8 f32 2>&1 | tee gpurun_out/r02_seed_sweep_chain64.log | tail -9
echo "== fixed-seed curve"; timeout 300 python scripts/swimmer_curve_gpu.py 40 swimmer f32 2>&1 | tail -3
timeout 300 python scripts/swimmer_curve_gpu.py 40 hopper f32 2>&1 | tail -2

#!/bin/bash
# tcgen05 32-wide update kernel: parity tests, then A/B against the FP32 tile kernel (variants/libb200rl_tile32.so)
mkdir -p gpurun_out
echo "== pytest update kernels"; timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fullsize.py tests/test_gpu_round2.py -q -m gpu --tb=short -p no:cacheprovider -x -k "not learning_curve" 2>&1 | tail -15
for wl in cartpole_vpg_65536x200 swimmer_trpo_16384x500; do
  for lib in rllab_b200/csrc/libb200rl.so rllab_b200/csrc/variants/libb200rl_tile32.so; do
    echo "== $wl $lib"
    timeout 300 python scripts/ab_lib.py $lib --workload $wl --steps 5 --warmup 3 --no-cpu-baseline --no-extra 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], {k:v['ms'] for k,v in d['kernels'].items()}, d['stats'])"
  done
done 2>&1 | tee gpurun_out/r02_ab_umma32.log

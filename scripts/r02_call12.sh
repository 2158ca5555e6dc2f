#!/bin/bash
# process_samples (cp.async staged scan) + umma32 tuning A/B (packed Gram x resident CTAs)
mkdir -p gpurun_out
echo "== pytest subset"; timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_round2.py tests/test_gpu_fullsize.py -q -m gpu --tb=short -p no:cacheprovider -x -k "not learning_curve" 2>&1 | tail -8
for wl in cartpole_vpg_65536x200 swimmer_trpo_16384x500; do
  for lib in rllab_b200/csrc/libb200rl.so rllab_b200/csrc/variants/libb200rl_p0b3.so rllab_b200/csrc/variants/libb200rl_p1b2.so rllab_b200/csrc/variants/libb200rl_p0b2.so; do
    echo "== $wl $lib"
    timeout 300 python scripts/ab_lib.py $lib --workload $wl --steps 5 --warmup 3 --no-cpu-baseline --no-extra 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], {k:v['ms'] for k,v in d['kernels'].items()})"
  done
done 2>&1 | tee gpurun_out/r02_ab_umma32_tuning.log

#!/bin/bash
# A/B of the experimental build variants against the default library:
#   cw = weights through the constant bank, ft = 5-instruction tanh, cwft = both.
# Build first (here or on the box):  for v in cw ft cwft; do make -C rllab_b200/csrc VARIANT=$v -j8; done
# and remove rllab_b200/csrc/variants/ from .gpurunignore so that the variant libraries travel.  Parity first, then timing.
for v in ${@:-cw ft cwft}; do
  V=$PWD/rllab_b200/csrc/variants/libb200rl_$v.so
  echo "== $v: parity"
  B200RL_LIB=$V python -m pytest tests/test_gpu_kernels.py tests/test_gpu_algos.py -q -m gpu 2>&1 | tail -6
done
for lib in default ${@:-cw ft cwft}; do
  V=$PWD/rllab_b200/csrc/variants/libb200rl_$lib.so; [ $lib = default ] && V=
  B200RL_LIB=$V python bench.py --steps 5 --warmup 3 --no-cpu-baseline | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('$lib', '%.3e'%d['value'], '%.2f'%d['ms_per_step'], {k:v['ms'] for k,v in d['kernels'].items()})"
done

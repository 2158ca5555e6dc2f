#!/bin/bash
# A/B of the experimental constant-bank-weights variant against the default library.
# Build first (here or on the box):  make -C rllab_b200/csrc VARIANT=cw -j8    and remove rllab_b200/csrc/variants/ from
# .gpurunignore so that the variant library travels.  Parity first, then timing.
V=$PWD/rllab_b200/csrc/variants/libb200rl_cw.so
B200RL_LIB=$V python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "rollout or loss_kl_grad_fvp or ragged" 2>&1 | tail -4
for lib in "" "$V"; do
B200RL_LIB=$lib python bench.py --steps 5 --warmup 3 --no-cpu-baseline | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('${lib:-default}', '%.3e'%d['value'], '%.2f'%d['ms_per_step'], {k:v['ms'] for k,v in d['kernels'].items()})"
done

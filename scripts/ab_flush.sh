#!/bin/bash
# check + timing of the 64-wide tiled-GEMM kernels after the two-phase flush / vectorised activation-cache I/O
python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "loss_kl_grad_fvp or f64_parity" 2>&1 | tail -3
python -m pytest tests/test_gpu_algos.py -q -m gpu -k "64 or hopper" 2>&1 | tail -3
B200RL_UPDATE_IMPL=gemm python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "loss_kl_grad_fvp" 2>&1 | tail -3
python bench.py --workload hopper_trpo_4096x500 --steps 5 --warmup 3 --no-cpu-baseline | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('hopper', '%.3e'%d['value'], '%.2f'%d['ms_per_step'], {k:v['ms'] for k,v in d['kernels'].items()})"

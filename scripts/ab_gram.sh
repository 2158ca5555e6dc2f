#!/bin/bash
# check + timing of the packed-FFMA2 Gram accumulation (tile and tiled-GEMM update kernels)
python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "loss_kl_grad_fvp or f64_parity or min_std" 2>&1 | tail -4
python -m pytest tests/test_gpu_algos.py -q -m gpu -k "trpo_step or vpg" 2>&1 | tail -3
for w in cartpole_vpg_65536x200 hopper_trpo_4096x500 swimmer_trpo_16384x500; do
python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('$w', '%.3e'%d['value'], '%.2f'%d['ms_per_step'], {k:v['ms'] for k,v in d['kernels'].items()})"
done

#!/bin/bash
# A/B of the 64-wide tiled-GEMM update kernels: 128 vs 256 threads per 128-sample tile (B200RL_GEMM_THREADS)
mkdir -p gpurun_out
B200RL_GEMM_THREADS=256 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "loss_kl_grad_fvp or lfb or golden or oracle_large" 2>&1 | tail -8
B200RL_GEMM_THREADS=256 python -m pytest tests/test_gpu_algos.py -q -m gpu -k "64" 2>&1 | tail -3
python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "loss_kl_grad_fvp and (hopper or swimmer)" 2>&1 | tail -3
for t in 128 256; do
B200RL_GEMM_THREADS=$t python bench.py --workload hopper_trpo_4096x500 --steps 4 --warmup 3 --no-cpu-baseline | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('threads=$t', '%.3e'%d['value'], d['ms_per_step'], {k:v['ms'] for k,v in d['kernels'].items()})"
done

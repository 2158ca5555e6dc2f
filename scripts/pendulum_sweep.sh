#!/bin/bash
# BASELINE.json configs[4]: Pendulum, VPG + LinearFeatureBaseline, roofline sweep N_envs in 2^12 .. 2^18 (per GPU)
for n in 4096 8192 16384 32768 65536 131072 262144; do
python bench.py --workload pendulum_vpg_262144x200 --lanes $n --steps 5 --warmup 3 --no-cpu-baseline | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); k=d['kernels']
print('| %d | %.3f | %.3e | %.3f | %.1f | %.3f | %.3f | %.3f | %.0f |' % ($n, d['ms_per_step'], d['value'], k['rollout']['ms'], k['rollout']['GBps'], k['grad']['ms'], k['loss_kl']['ms'], k['process_samples']['ms'], k['process_samples']['GBps']))"
done

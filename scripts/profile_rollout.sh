#!/bin/bash
# ncu --set full of the rollout kernel for the planar workloads (one launch each)
mkdir -p gpurun_out
for w in swimmer_trpo_16384x500 hopper_trpo_4096x500; do
ncu --set full --clock-control none --import-source on -k regex:'rollout_kernel' -s 1 -c 1 -o gpurun_out/prof_rollout_$w -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --workload $w > gpurun_out/ncu_rollout_$w.log 2>&1
done
ls -la gpurun_out

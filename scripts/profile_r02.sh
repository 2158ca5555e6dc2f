#!/bin/bash
# Round-2 profiles (run on the GPU box under gpurun):  bash scripts/profile_r02.sh
#  1) launch list (per-kernel device time) of the default bench command
#  2) ncu --set full of the hot kernels of cfg2 (tcgen05 gradient and loss passes, rollout, process_samples, baseline Gram)
#  3) ncu --set full of the tcgen05 gradient / Fisher-vector kernels on the Hopper (64,64) and Swimmer (32,32) workloads
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/bench_under_ncu_r02.log 2>&1
ncu --set full --clock-control none --import-source on \
    -k regex:'update_umma32_kernel|loss_thread_kernel|rollout_kernel|gae_scan_kernel|lfb_predict_kernel|lfb_gram' \
    -s 6 -c 6 -o gpurun_out/prof_r02_cfg2 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra \
    > gpurun_out/ncu_full_r02_cfg2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'update_umma64_kernel' -s 2 -c 3 \
    -o gpurun_out/prof_r02_hopper -f python bench.py --workload hopper_trpo_4096x500 --steps 1 --warmup 1 \
    --no-cpu-baseline --no-extra > gpurun_out/ncu_full_r02_hopper.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'update_umma32_kernel' -s 3 -c 2 \
    -o gpurun_out/prof_r02_swimmer -f python bench.py --workload swimmer_trpo_16384x500 --steps 1 --warmup 1 \
    --no-cpu-baseline --no-extra > gpurun_out/ncu_full_r02_swimmer.log 2>&1
ls -la gpurun_out | tail -6

#!/bin/bash
# 2 GPUs: peer worker directly (full traceback), then the 1-GPU suite + bench + ncu of the tcgen05 32-wide gradient kernel
mkdir -p gpurun_out
echo "== peer worker"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 tests/peer_worker.py > gpurun_out/r02_peer_worker.log 2>&1; grep -v "^\[W\|frame #" gpurun_out/r02_peer_worker.log | tail -40
echo "== pytest gpu"; timeout 1500 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider --deselect tests/test_gpu_round2.py::test_swimmer_learning_curve_matches_oracle --deselect tests/test_gpu_round2.py::test_hopper_learning_curve_matches_oracle > gpurun_out/r02_pytest10.log 2>&1; tail -25 gpurun_out/r02_pytest10.log
echo "== bench"; timeout 600 python bench.py 2> gpurun_out/r02_bench10.err | tail -1 > gpurun_out/r02_bench10.json; python -c "
import json
d=json.load(open('gpurun_out/r02_bench10.json')); print(d['ms_per_step'], d['value'], d['e2e'], {k:v['ms'] for k,v in d['kernels'].items()}, d['roofline']['frac'], d['roofline_hbm'], d['cpu_baseline'], d['extra'])"
tail -3 gpurun_out/r02_bench10.err
echo "== ncu grad umma32"
ncu --set full --clock-control none --import-source on -k regex:'update_umma32_kernel|gae_scan_kernel|lfb_predict_kernel' -s 6 -c 3 -o gpurun_out/prof_r02b_cfg2 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/ncu_full_r02b_cfg2.log 2>&1
ls -la gpurun_out | tail -5

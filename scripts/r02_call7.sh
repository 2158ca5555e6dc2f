#!/bin/bash
# round-2 (session 2) call: full GPU suite, default bench, launch list + ncu --set full captures
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider -x --deselect tests/test_gpu_round2.py::test_swimmer_learning_curve_matches_oracle --deselect tests/test_gpu_round2.py::test_hopper_learning_curve_matches_oracle > gpurun_out/r02_pytest7.log 2>&1; tail -15 gpurun_out/r02_pytest7.log
echo "== bench"; timeout 600 python bench.py 2> gpurun_out/r02_bench7.err | tail -1 | tee gpurun_out/r02_bench7.json
tail -3 gpurun_out/r02_bench7.err
echo "== profiles"; timeout 900 bash scripts/profile_r02.sh

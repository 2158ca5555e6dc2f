#!/bin/bash
# A/B: bench the default library and every rllab_b200/csrc/variants/*.so on the headline workload
python bench.py --steps 5 --warmup 3 --no-cpu-baseline | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('base', '%.3e'%d['value'], d['ms_per_step'], {k:v['ms'] for k,v in d['kernels'].items()})"
for v in rllab_b200/csrc/variants/*.so; do
B200RL_LIB=$PWD/$v python bench.py --steps 5 --warmup 3 --no-cpu-baseline | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.readline()); print('$v', '%.3e'%d['value'], d['ms_per_step'], {k:v['ms'] for k,v in d['kernels'].items()})"
done

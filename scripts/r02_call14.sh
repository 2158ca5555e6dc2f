#!/bin/bash
# 4 GPUs: peer-memory exchange with more than two ranks (worker + bench at cfg2 with both transports)
mkdir -p gpurun_out
echo "== peer worker x4"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29721 tests/peer_worker.py > gpurun_out/r02_peer_worker4.log 2>&1; grep -v "^\[W\|frame #\|OMP_NUM\|\*\*\*\*" gpurun_out/r02_peer_worker4.log | tail -12
for peer in 1 0; do
echo "== bench 4 GPUs B200RL_PEER=$peer"
B200RL_PEER=$peer timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29722 bench.py --gpus 4 --steps 10 --warmup 3 2> gpurun_out/r02_bench14_$peer.err | tail -1 | tee gpurun_out/r02_bench14_peer$peer.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['collectives_per_step'], d.get('peer_exchanges_per_step'), d.get('transport'), d['shard_check'], d['replicas_identical'], {k:(v['ms_per_step']) for k,v in d['extra']['workloads'].items()})"
grep -v "^\[W\|frame #\|OMP_NUM\|\*\*\*\*" gpurun_out/r02_bench14_$peer.err | tail -3
done

#!/bin/bash
# 2 GPUs: peer-memory exchange tests, 2-GPU bench (cfg2) with and without the peer transport
mkdir -p gpurun_out
nvidia-smi topo -m 2>&1 | head -8
echo "== pytest peer"; timeout 900 python -m pytest tests/test_gpu_peer.py -q -m gpu --tb=short -p no:cacheprovider -x 2>&1 | tail -30
for peer in 1 0; do
echo "== bench 2 GPUs B200RL_PEER=$peer"
B200RL_PEER=$peer timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus 2 --steps 10 --warmup 3 --no-extra 2> gpurun_out/r02_bench9_$peer.err | tail -1 | tee gpurun_out/r02_bench9_peer$peer.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['collectives_per_step'], d.get('peer_exchanges_per_step'), d.get('transport'), d['shard_check'], d['replicas_identical'])"
tail -3 gpurun_out/r02_bench9_$peer.err
done
echo "== pytest kernels (1 GPU)"; timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fullsize.py -q -m gpu --tb=short -p no:cacheprovider -x 2>&1 | tail -5

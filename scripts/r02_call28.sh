#!/bin/bash
# 2 GPUs: peer tests with the final library (timeout counter, 32-slice fused finalize)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_peer.py -q -m gpu --tb=short -p no:cacheprovider 2>&1 | tail -6

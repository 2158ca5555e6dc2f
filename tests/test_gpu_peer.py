"""Multi-GPU (needs >= 2 GPUs on the box; skipped otherwise): the peer-memory exchange of rllab_b200/csrc/peer.cuh --
stand-alone and fused into the update passes -- against NCCL and against the single-rank pass, and bench.py's own
rank-count invariance check on 2 GPUs."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _n_gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _torchrun(n, script, *args, port=29631):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr",
           "127.0.0.1", "--master-port", str(port), script] + list(args)
    return subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)


@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs")
def test_peer_exchange_matches_nccl_and_single_rank():
    out = _torchrun(2, os.path.join(ROOT, "tests", "peer_worker.py"))
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "PEER_OK" in out.stdout and "nccl=0" in out.stdout, out.stdout[-2000:]


@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs")
def test_bench_two_gpus_reproduces_single_process_run():
    out = _torchrun(2, os.path.join(ROOT, "bench.py"), "--gpus", "2", "--lanes", "8192", "--steps", "3", "--warmup", "3",
                    "--no-cpu-baseline", port=29632)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["n_gpus"] == 2 and line["replicas_identical"] is True
    assert line["shard_check"]["max_rel_diff_vs_single_process"] < 1e-7
    assert line["collectives_per_step"] == 0 and line["peer_exchanges_per_step"] >= 3       # no NCCL on the critical path
    assert line["transport"].startswith("peer-memory")

"""CPU, world_size 2, gloo: the multi-GPU host logic (rllab_b200/parallel.py) -- contiguous lane sharding, one
all-reduce per reduction vector, replicated deterministic update -- reproduces the single-process result."""
import json
import os
import subprocess
import sys

import numpy as np

from oracle import envs as E, policy as P, sampler as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_is_a_contiguous_partition():
    from rllab_b200.parallel import Comm

    class FakeComm(Comm):
        def __init__(self, rank, world):
            self.rank, self.world_size, self.active = rank, world, world > 1
    for n_total, world in ((65536, 8), (10, 4), (7, 2), (5, 5)):
        spans = [FakeComm(r, world).shard(n_total) for r in range(world)]
        assert sum(n for n, _ in spans) == n_total
        pos = 0
        for n, lane0 in spans:
            assert lane0 == pos
            pos += n
        assert max(n for n, _ in spans) - min(n for n, _ in spans) <= 1


def test_two_rank_gloo_run_matches_single_process(tmp_path):
    env_vars = dict(os.environ, PYTHONPATH=ROOT, OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29731", os.path.join(ROOT, "tests", "dist_worker.py"), str(tmp_path)]
    res = subprocess.run(cmd, env=env_vars, capture_output=True, text=True, timeout=240)
    assert res.returncode == 0, res.stderr[-2000:]
    ranks = [json.load(open(os.path.join(str(tmp_path), "rank%d.json" % r))) for r in range(2)]
    # single-process reference on all lanes
    N, T, mpl = 24, 30, 11
    env = E.make("cartpole")
    dims = P.Dims(env.O, (8, 8), env.A)
    theta = P.init_params(dims, np.random.RandomState(0))
    rng = np.random.RandomState(1)
    eps = rng.randn(T, env.A, N)
    rr = rng.rand(T + 1, env.K, N)
    traj = S.rollout_lanes(env, theta, dims, N, T, mpl, eps, rr)
    w = np.random.RandomState(2).randn(2 * env.O + 4) * 0.1
    full = S.process_samples_lanes(traj, w, 0.99, 0.97, center_adv=True, drop_cut=True)
    valid = full["valid"]
    assert 0 < (~valid).sum() < valid.size                     # the case exercises dropped (cut) paths
    adv_c = np.concatenate([np.array(r["adv_c"]) for r in ranks], axis=1)
    np.testing.assert_allclose(adv_c, full["adv"], rtol=1e-10, atol=1e-12)
    coeffs = S.lfb_fit_lanes(traj["obs"], traj["tstep"], full["ret"], valid=valid)
    batch = S.batch_from_traj(traj, full["adv"], valid)
    g = P.grad_surr(theta, batch, dims, "trpo")
    tri = [P.surr_loss_trpo(theta, batch, dims)] + list(P.kl_stats(theta + 1e-3, batch, dims))
    x = np.random.RandomState(3).randn(dims.P)
    Hx = P.fvp(theta, batch, x, dims, 1e-5)
    for r in ranks:                      # every rank ends with the same replicated values
        np.testing.assert_allclose(r["coeffs"], coeffs, rtol=1e-6, atol=1e-8)
        np.testing.assert_allclose(r["g"], g, rtol=1e-10, atol=1e-14)
        np.testing.assert_allclose(r["Hx"], Hx, rtol=1e-10, atol=1e-14)
        np.testing.assert_allclose(r["mx"], [full["adv_raw"][valid].max(), -full["adv_raw"][valid].min()], rtol=1e-12)
        np.testing.assert_allclose(r["tri"], tri, rtol=1e-10, atol=1e-14)       # sums and the maximum in one message
        assert r["n_collectives"] == 3                                         # statistics+fit, gradient+triple, FVP
    assert ranks[0]["g"] == ranks[1]["g"] and ranks[0]["coeffs"] == ranks[1]["coeffs"]     # bit-identical replicas
    assert (ranks[0]["lane0"], ranks[0]["n_local"], ranks[1]["lane0"], ranks[1]["n_local"]) == (0, 12, 12, 12)

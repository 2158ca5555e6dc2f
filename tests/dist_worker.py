"""Worker for tests/test_dist_cpu.py: world_size-2 gloo run of the host-side multi-GPU logic (lane sharding,
reduction-vector all-reduce, identical replicated update on every rank) with the CPU oracle standing in for the
device kernels.  Launched by torch.distributed.run; writes one JSON line per rank to the directory in argv[1]."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import envs as E, policy as P, sampler as S, optim as OPT   # noqa: E402
from rllab_b200.parallel import Comm                                      # noqa: E402


def main():
    outdir = sys.argv[1]
    comm = Comm(backend="gloo")
    assert comm.active and comm.world_size == 2
    N, T, mpl = 24, 30, 11
    env = E.make("cartpole")
    dims = P.Dims(env.O, (8, 8), env.A)
    theta = P.init_params(dims, np.random.RandomState(0))
    rng = np.random.RandomState(1)
    eps = rng.randn(T, env.A, N)
    rr = rng.rand(T + 1, env.K, N)
    n_local, lane0 = comm.shard(N)
    sl = slice(lane0, lane0 + n_local)
    # each rank rolls out only its contiguous lane block (noise is keyed by the GLOBAL lane index)
    traj = S.rollout_lanes(env, theta, dims, n_local, T, mpl, eps[:, :, sl], rr[:, :, sl])
    w = np.random.RandomState(2).randn(2 * env.O + 4) * 0.1
    # --- process_samples (whole paths only: cut paths dropped) + baseline normal equations: ONE mixed collective
    # [adv sums (3) | normal equations | maxima (2)], exactly the layout LaneBatch.red uses on the device
    loc = S.process_samples_lanes(traj, w, 0.99, 0.97, center_adv=False, drop_cut=True)
    valid = loc["valid"]
    adv = loc["adv_raw"]
    F = S.lfb_features_lanes(traj["obs"], traj["tstep"])
    d = F.shape[0]
    Fm = np.concatenate([F.reshape(d, -1), loc["ret"].reshape(1, -1)])[:, valid.reshape(-1)]
    G_loc = (Fm @ Fm.T).reshape(-1)
    red = torch.tensor(np.concatenate([[adv[valid].sum(), (adv[valid] ** 2).sum(), float(valid.sum())], G_loc,
                                       [adv[valid].max(), -adv[valid].min()]]), dtype=torch.float64)
    comm.all_reduce_mixed(red, 3 + G_loc.size)
    sums, Gn, mx = red[:3], red[3:3 + G_loc.size].numpy().reshape(d + 1, d + 1), red[3 + G_loc.size:]
    mean = sums[0] / sums[2]
    std = torch.sqrt(sums[1] / sums[2] - mean ** 2)
    adv_c = np.where(valid, (adv - float(mean)) / (float(std) + 1e-8), 0.0)
    coeffs = S.lfb_fit_normal(Gn[:-1, :-1], Gn[:-1, -1])
    # --- gradient (+ loss / KL triple in the same message) and FVP: local sums over the valid samples divided by the
    # GLOBAL valid count, then one collective == global mean
    batch = S.batch_from_traj(traj, adv_c, valid)
    B_local, count = float(valid.sum()), float(sums[2])
    kl_mean, kl_max = P.kl_stats(theta + 1e-3, batch, dims)
    gl = torch.tensor(np.concatenate([P.grad_surr(theta, batch, dims, "trpo") * (B_local / count),
                                      [P.surr_loss_trpo(theta, batch, dims) * (B_local / count),
                                       kl_mean * (B_local / count), kl_max]]))
    comm.all_reduce_mixed(gl, dims.P + 2)
    g, tri = gl[:dims.P], gl[dims.P:]
    x = np.random.RandomState(3).randn(dims.P)
    Hx = torch.tensor((P.fvp(theta, batch, x, dims, 0.0) * (B_local / count)) + 1e-5 * x / comm.world_size)
    comm.all_reduce_mixed(Hx, Hx.numel())
    comm.barrier()
    out = dict(rank=comm.rank, lane0=lane0, n_local=n_local, adv_c=adv_c.tolist(), coeffs=coeffs.tolist(),
               g=g.numpy().tolist(), Hx=Hx.numpy().tolist(), mx=mx.numpy().tolist(), tri=tri.numpy().tolist(),
               n_collectives=comm.n_collectives)
    with open(os.path.join(outdir, "rank%d.json" % comm.rank), "w") as f:
        json.dump(out, f)
    comm.close()


if __name__ == "__main__":
    main()

"""Worker of tests/test_gpu_peer.py (one process per GPU under torchrun): the peer-memory exchange of csrc/peer.cuh against
NCCL on the same vectors -- stand-alone mixed all-reduce (sums + maxima, many back-to-back calls of changing size, so that
the sequence-number / double-buffer protocol is exercised without host synchronisation in between) and the exchange fused
into the finalize kernel of an update pass."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from rllab_b200 import ops
    from rllab_b200 import _lib as L
    from rllab_b200.parallel import Comm
    comm = Comm()
    dev = torch.device("cuda", comm.local_rank)
    assert comm.active and comm.peer, "peer-memory transport did not come up"
    dist = comm.dist
    rng = np.random.RandomState(100 + comm.rank)
    # ---- stand-alone exchange, 200 calls queued without a host sync
    outs, refs = [], []
    for it in range(200):
        n = int(rng.randint(1, 6000)) if it % 7 else 8192
        n = comm_same_int(dist, n, dev)
        n_sum = (n * 2) // 3 if it % 3 else n
        t = torch.tensor(rng.randn(n), dtype=torch.float64, device=dev)
        ref = t.clone()
        comm.all_reduce_mixed(t, n_sum)
        outs.append(t)
        refs.append((ref, n_sum))
    torch.cuda.synchronize()
    for t, (ref, n_sum) in zip(outs, refs):
        g = [torch.empty_like(ref) for _ in range(comm.world_size)]
        dist.all_gather(g, ref)
        g = torch.stack(g)
        want = torch.cat([rank_order_sum(g[:, :n_sum]), g[:, n_sum:].max(0).values])
        assert torch.equal(t, want), (t - want).abs().max()
    # ---- fused into an update pass: sharded loss / gradient / FVP == the same pass over the whole batch on one rank
    from oracle import envs as E, policy as P
    env = E.make("cartpole", np.float32)
    dims = P.Dims(env.O, (32, 32), env.A)
    theta = P.init_params(dims, np.random.RandomState(1))
    th32 = torch.tensor(theta, dtype=torch.float32, device=dev)
    N, T, W = 512, 40, comm.world_size
    dd = (env.O, 32, 32, env.A)

    def make_batch(n, lane0, n_total):
        b = ops.LaneBatch(env.O, env.A, n, T, dev)
        ops.rollout(L.ENV_CARTPOLE, th32, 32, 32, 1e-6, b, T, None, None, 3, 0, lane0)
        ops.process_samples(b, None, 0.99, 1.0)
        return b

    full = make_batch(N * W, 0, N * W)
    ops.center_advantages(full, True, False)
    mine = make_batch(N, comm.rank * N, N * W)
    mine.adv.copy_(full.adv.view(T, N * W)[:, comm.rank * N:(comm.rank + 1) * N])   # globally centred
    mine.B_global = N * W * T
    P_ = dims.P
    th2 = torch.tensor(theta + 0.02 * np.random.RandomState(9).randn(P_), dtype=torch.float32, device=dev)
    x = torch.tensor(np.random.RandomState(5).randn(P_), dtype=torch.float64, device=dev)
    res = {}
    for name, b, fuse in (("full", full, False), ("shard", mine, True)):
        g, tri, Hx, out = (torch.zeros(k, dtype=torch.float64, device=dev) for k in (P_, 3, P_, 3))
        ops.grad(L.LOSS_TRPO, th2, dd, 1e-6, b, g, tri, b.hcache(32, 32), fuse=fuse)
        ops.fvp(th2, dd, 1e-6, b, x, 1e-5, 1.0 / (W if fuse else 1), Hx, b.hcache(32, 32), fuse=fuse)
        ops.loss_kl(L.LOSS_TRPO, th2, dd, 1e-6, b, out, fuse=fuse)
        res[name] = [v.cpu().numpy() for v in (g, tri, Hx, out)]
    for a, b_ in zip(res["full"], res["shard"]):
        # tiles group different samples in the sharded layout: float32 per-tile partial sums differ in the last bits
        np.testing.assert_allclose(b_, a, rtol=0, atol=2e-6 * np.abs(a).max())
    # every rank holds bit-identical results
    for v in res["shard"]:
        t = torch.tensor(v, device=dev)
        g = [torch.empty_like(t) for _ in range(W)]
        dist.all_gather(g, t)
        assert all(torch.equal(g[0], q) for q in g)
    if comm.rank == 0:
        print("PEER_OK exchanges=%d nccl=%d" % (comm.n_peer_exchanges, comm.n_collectives))
    comm.close()


def comm_same_int(dist, n, dev):
    t = torch.tensor([n], dtype=torch.int64, device=dev)
    dist.broadcast(t, 0)
    return int(t.item())


def rank_order_sum(g):
    acc = g[0].clone()
    for r in range(1, g.shape[0]):
        acc += g[r]
    return acc


if __name__ == "__main__":
    main()

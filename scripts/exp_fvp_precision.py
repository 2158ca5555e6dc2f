"""Experiment (not product code): which part of a float32 Fisher-vector product costs TRPO its learning speed on Swimmer?
The CG solve of the swimmer_curve_gpu.py configuration is run with a torch implementation of the product whose tangent-
forward and backward halves can be evaluated in float32 or float64 independently (gradient, loss and line search stay on the
float32 kernels).  Usage: python scripts/exp_fvp_precision.py <fwd f32|f64> <bwd f32|f64> <x f32|f64> [seed] [n_itr]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def make_hx(opt, fwd, bwd, xdt):
    DT = dict(f32=torch.float32, f64=torch.float64)
    F, Bk, XD = DT[fwd], DT[bwd], DT[xdt]

    def _make_Hx(batch, b, hcache, tiles):
        pol = opt._target
        O, h1n, h2n, A = pol.dims
        th = pol.theta32.double()
        k = 0
        parts = []
        for shp in [(O, h1n), (h1n,), (h1n, h2n), (h2n,), (h2n, A), (A,), (A,)]:
            n = int(np.prod(shp))
            parts.append((k, n, shp))
            k += n
        get = lambda v, i: v[parts[i][0]:parts[i][0] + parts[i][1]].reshape(parts[i][2])
        X = batch.obs.reshape(O, -1).t().double()
        W0, b0, W1, b1, Wo, bo = (get(th, i) for i in range(6))
        H1 = torch.tanh(X @ W0 + b0)
        H2 = torch.tanh(H1 @ W1 + b1)
        ls = torch.clamp(get(th, 6), min=float(np.log(pol.min_std)))
        s2 = torch.exp(2 * ls)
        Mmu = 2.0 / (2.0 * s2 + 1e-8)
        Ml = 4 * s2 * (2 * s2 - 1e-8) / (2 * s2 + 1e-8) ** 2
        Bn = X.shape[0]

        def Hx(vec, out):
            v = vec.to(XD).double()
            V0, vb0, V1, vb1, Vo, vbo = (get(v, i) for i in range(6))
            c = lambda t: t.to(F)
            t1 = (c(X) @ c(V0) + c(vb0)) * (1 - c(H1) ** 2)
            t2 = (t1 @ c(W1) + c(H1) @ c(V1) + c(vb1)) * (1 - c(H2) ** 2)
            md = t2 @ c(Wo) + c(H2) @ c(Vo) + c(vbo)
            dmu = (md.double() * Mmu / Bn).to(Bk)
            cb = lambda t: t.to(Bk)
            d2 = (dmu @ cb(Wo).t()) * (1 - cb(H2) ** 2)
            d1 = (d2 @ cb(W1).t()) * (1 - cb(H1) ** 2)
            res = torch.cat([(cb(X).t() @ d1).reshape(-1), d1.sum(0), (cb(H1).t() @ d2).reshape(-1), d2.sum(0),
                             (cb(H2).t() @ dmu).reshape(-1), dmu.sum(0), torch.zeros(A, dtype=Bk, device=X.device)]).double()
            res[-A:] = Ml * get(vec, 6)
            out.copy_(res + opt._reg_coeff * vec)
        return Hx
    return _make_Hx


def main(fwd, bwd, xdt, seed, n_itr):
    from rllab_b200.algos.trpo import TRPO
    from rllab_b200.baselines.linear_feature_baseline import LinearFeatureBaseline
    from rllab_b200.envs.mujoco.swimmer_env import SwimmerEnv
    from rllab_b200.envs.normalized_env import normalize
    from rllab_b200.misc import logger
    from rllab_b200.policies.gaussian_mlp_policy import GaussianMLPPolicy
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "oracle_swimmer_trpo_curve.json")))
    cfg = gold["config"]
    logger.set_quiet(True)
    env = normalize(SwimmerEnv())
    policy = GaussianMLPPolicy(env.spec, hidden_sizes=tuple(cfg["hidden"]), seed=cfg["policy_seed"])
    algo = TRPO(env=env, policy=policy, baseline=LinearFeatureBaseline(env.spec), batch_size=cfg["lanes"] * cfg["horizon"],
                max_path_length=cfg["horizon"], n_itr=n_itr, discount=cfg["discount"], gae_lambda=cfg["gae_lambda"],
                step_size=cfg["step_size"], optimizer_args=dict(cg_iters=cfg["cg_iters"]),
                sampler_args=dict(n_envs=cfg["lanes"], seed=seed))
    algo.start_worker()
    algo.init_opt()
    algo.optimizer._make_Hx = make_hx(algo.optimizer, fwd, bwd, xdt)
    rets = []
    for itr in range(n_itr):
        algo.train_itr(itr)
        rets.append(logger.get_last_table()["AverageReturn"])
    print("fwd %s bwd %s x %s seed %d: last-5 mean %.3f  (itr 9: %.2f, 19: %.2f, 29: %.2f, 39: %.2f)" %
          (fwd, bwd, xdt, seed, np.mean(rets[-5:]), rets[9], rets[19], rets[29], rets[-1]), flush=True)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 7,
         int(sys.argv[5]) if len(sys.argv) > 5 else 40)

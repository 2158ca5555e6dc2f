"""Experiment (not product code): how accurate is the float32 Fisher-vector product on the Swimmer batch, and is the
computed operator symmetric?  Compares b200rl_fvp (float32 kernels) with b200rl_update_f64 (mode 2) on the same batch for
random directions, the gradient, and the directions a float64 CG solve generates; prints relative errors, the symmetry
defect q.Ap - p.Aq, and the same for the two-word (hi + lo) evaluation."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from rllab_b200 import _lib as L, ops
    from rllab_b200.algos.trpo import TRPO
    from rllab_b200.baselines.linear_feature_baseline import LinearFeatureBaseline
    from rllab_b200.envs.mujoco.swimmer_env import SwimmerEnv
    from rllab_b200.envs.normalized_env import normalize
    from rllab_b200.misc import logger
    from rllab_b200.policies.gaussian_mlp_policy import GaussianMLPPolicy
    cfg = json.load(open(os.path.join(ROOT, "tests", "golden", "oracle_swimmer_trpo_curve.json")))["config"]
    logger.set_quiet(True)
    env = normalize(SwimmerEnv())
    policy = GaussianMLPPolicy(env.spec, hidden_sizes=tuple(cfg["hidden"]), seed=cfg["policy_seed"])
    algo = TRPO(env=env, policy=policy, baseline=LinearFeatureBaseline(env.spec), batch_size=cfg["lanes"] * cfg["horizon"],
                max_path_length=cfg["horizon"], n_itr=40, discount=cfg["discount"], gae_lambda=cfg["gae_lambda"],
                step_size=cfg["step_size"], optimizer_args=dict(cg_iters=10, precision="f64"),
                sampler_args=dict(n_envs=cfg["lanes"], seed=7))
    algo.start_worker()
    algo.init_opt()
    for n_warm in (0, 15, 15):                       # look at iteration 0, 15 and 30 of a float64-CG run
        for itr in range(n_warm):
            algo.train_itr(algo.current_itr if hasattr(algo, "current_itr") and algo.current_itr else itr)
        itr_now = getattr(algo, "current_itr", 0) or 0
        paths = algo.sampler.obtain_samples(itr_now)
        sd = algo.sampler.process_samples(itr_now, paths)
        b = sd["lane_batch"] if isinstance(sd, dict) else sd.lane_batch
        pol = policy
        P = pol.n_params
        dev = b.device
        z = lambda: torch.zeros(P, dtype=torch.float64, device=dev)
        g, hc = z(), b.hcache(pol.h1, pol.h2)
        ops.grad(L.LOSS_TRPO, pol.theta32, pol.dims, pol.min_std, b, g, None, hc)
        th64 = pol.theta32.double()                   # same parameters for both operators

        def A32(p):
            out = z()
            ops.fvp(pol.theta32, pol.dims, pol.min_std, b, p, 1e-5, 1.0, out, hc)
            return out

        def A32hl(p):
            hi = p.float().double()
            return A32(hi) + A32(p - hi)

        def A64(p):
            out = z()
            ops.update_f64(2, L.LOSS_TRPO, th64, pol.dims, pol.min_std, b, p, 1e-5, 1.0, out, None)
            return out

        # float64 CG directions
        x, r, p = z(), g.clone(), g.clone()
        dirs = []
        rr = r.dot(r)
        for k in range(10):
            dirs.append(p.clone())
            Ap = A64(p)
            v = rr / p.dot(Ap)
            x += v * p
            r -= v * Ap
            rr2 = r.dot(r)
            p = r + (rr2 / rr) * p
            rr = rr2
        rng = np.random.RandomState(0)
        rnd = [torch.tensor(rng.randn(P), dtype=torch.float64, device=dev) for _ in range(2)]
        print("---- iteration %d   |g| %.3e   policy std %.3f" % (itr_now, float(g.norm()), float(pol.theta32[-1].exp())))
        for name, p in [("random0", rnd[0]), ("random1", rnd[1])] + [("cg dir %d" % k, d) for k, d in enumerate(dirs)]:
            a64, a32, ahl = A64(p), A32(p), A32hl(p)
            rq = float(p.dot(a64) / p.dot(p))                    # Rayleigh quotient of the direction
            e32 = float((a32 - a64).norm() / a64.norm())
            ehl = float((ahl - a64).norm() / a64.norm())
            q = rnd[1] if name != "random1" else rnd[0]
            asym32 = float((q.dot(a32) - p.dot(A32(q))) / (p.norm() * q.norm()))
            asym64 = float((q.dot(a64) - p.dot(A64(q))) / (p.norm() * q.norm()))
            print("%-9s rayleigh %.3e  rel err f32 %.2e  hi+lo %.2e   err/(reg|p|) %.2e   asym f32 %.2e f64 %.2e" %
                  (name, rq, e32, ehl, float((a32 - a64).norm() / (1e-5 * p.norm())), asym32, asym64), flush=True)


if __name__ == "__main__":
    main()

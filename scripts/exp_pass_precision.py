"""Experiment (not product code): which float32 pass costs TRPO its learning speed on Swimmer?  The optimizer's three passes
(gradient, Fisher-vector product, loss/KL of the line search) are switched between the float32 kernels and the float64 parity
kernels independently.  Usage: python scripts/exp_pass_precision.py <grad f32|f64> <fvp f32|f64> <loss f32|f64> [seed] [n_itr]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(g, h, l, seed, n_itr):
    from rllab_b200 import ops
    from rllab_b200.algos.trpo import TRPO
    from rllab_b200.baselines.linear_feature_baseline import LinearFeatureBaseline
    from rllab_b200.envs.mujoco.swimmer_env import SwimmerEnv
    from rllab_b200.envs.normalized_env import normalize
    from rllab_b200.misc import logger
    from rllab_b200.optimizers.conjugate_gradient_optimizer import ConjugateGradientOptimizer
    from rllab_b200.policies.gaussian_mlp_policy import GaussianMLPPolicy

    class Mixed(ConjugateGradientOptimizer):
        def _eval(self, batch, want_grad=False):
            self._f64 = (g if want_grad else l) == "f64"
            self._use_hcache = not self._f64 and h == "f32"
            try:
                return ConjugateGradientOptimizer._eval(self, batch, want_grad)
            finally:
                self._f64 = False

        def _make_Hx(self, batch, b, hcache, tiles):
            self._f64 = h == "f64"
            try:
                if h == "f32" and hcache is None and g == "f64":      # the f64 gradient pass wrote no activation cache
                    pol = self._target
                    hcache = batch.hcache(pol.h1, pol.h2)
                    ops.grad(self._loss_kind, pol.theta32, pol.dims, pol.min_std, batch, b["tmp"], None, hcache)
                inner = ConjugateGradientOptimizer._make_Hx(self, batch, b, hcache, tiles)
            finally:
                f64 = self._f64
                self._f64 = False

            def Hx(vec, out):
                self._f64 = f64
                try:
                    inner(vec, out)
                finally:
                    self._f64 = False
            return Hx

    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "oracle_swimmer_trpo_curve.json")))
    cfg = gold["config"]
    logger.set_quiet(True)
    env = normalize(SwimmerEnv())
    policy = GaussianMLPPolicy(env.spec, hidden_sizes=tuple(cfg["hidden"]), seed=cfg["policy_seed"])
    algo = TRPO(env=env, policy=policy, baseline=LinearFeatureBaseline(env.spec), batch_size=cfg["lanes"] * cfg["horizon"],
                max_path_length=cfg["horizon"], n_itr=n_itr, discount=cfg["discount"], gae_lambda=cfg["gae_lambda"],
                step_size=cfg["step_size"], optimizer=Mixed(cg_iters=cfg["cg_iters"]),
                sampler_args=dict(n_envs=cfg["lanes"], seed=seed))
    algo.start_worker()
    algo.init_opt()
    rets, bt = [], []
    for itr in range(n_itr):
        algo.train_itr(itr)
        rets.append(logger.get_last_table()["AverageReturn"])
        bt.append(algo.optimizer.last_info["n_iter"])
    print("grad %s fvp %s loss %s seed %d: last-5 mean %.3f  (itr 9: %.2f, 19: %.2f, 29: %.2f, 39: %.2f) mean backtracks %.2f" %
          (g, h, l, seed, np.mean(rets[-5:]), rets[9], rets[19], rets[29], rets[-1], np.mean(bt)), flush=True)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 7,
         int(sys.argv[5]) if len(sys.argv) > 5 else 40)

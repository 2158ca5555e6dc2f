"""Experiment (not product code): does a CONSISTENT float32 Fisher-vector product restore TRPO's learning speed on Swimmer?
The float32 kernels round the CG direction p to float32 on entry, so CG sees z = H fl32(p) instead of H p -- a systematic
error of 6e-8 |H| |p| that the small eigen-directions (reg 1e-5) cannot absorb.  Variants:
  base   the shipped float32 path
  roundp CG keeps p float32-representable (ConjugateGradientOptimizer(cg_direction_f32=True)): z = H p exactly for that p
  hilo   p = p_hi + p_lo (both float32-representable), z = H p_hi + H p_lo by linearity (two kernel calls)
Usage: python scripts/exp_fvp_hilo.py <base|roundp|hilo> [seed] [n_itr]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(variant, seed, n_itr, env_name="swimmer"):
    import torch
    from rllab_b200.algos.trpo import TRPO
    from rllab_b200.baselines.linear_feature_baseline import LinearFeatureBaseline
    from rllab_b200.envs.mujoco.hopper_env import HopperEnv
    from rllab_b200.envs.mujoco.swimmer_env import SwimmerEnv
    from rllab_b200.envs.normalized_env import normalize
    from rllab_b200.misc import logger
    from rllab_b200.optimizers.conjugate_gradient_optimizer import ConjugateGradientOptimizer
    from rllab_b200.policies.gaussian_mlp_policy import GaussianMLPPolicy

    class HiLo(ConjugateGradientOptimizer):
        def _make_Hx(self, batch, b, hcache, tiles):
            inner = ConjugateGradientOptimizer._make_Hx(self, batch, b, hcache, tiles)
            lo_out = torch.zeros_like(b["z"])

            def Hx(vec, out):
                hi = vec.float().double()
                lo = vec - hi
                inner(hi, out)
                inner(lo, lo_out)
                out.add_(lo_out)
            return Hx

    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "oracle_%s_trpo_curve.json" % env_name)))
    cfg = gold["config"]
    logger.set_quiet(True)
    env = normalize(HopperEnv() if env_name == "hopper" else SwimmerEnv())
    policy = GaussianMLPPolicy(env.spec, hidden_sizes=tuple(cfg["hidden"]), seed=cfg["policy_seed"])
    opt = dict(base=lambda: ConjugateGradientOptimizer(cg_iters=cfg["cg_iters"]),
               roundp=lambda: ConjugateGradientOptimizer(cg_iters=cfg["cg_iters"], cg_direction_f32=True),
               hilo=lambda: HiLo(cg_iters=cfg["cg_iters"]))[variant]()
    algo = TRPO(env=env, policy=policy, baseline=LinearFeatureBaseline(env.spec), batch_size=cfg["lanes"] * cfg["horizon"],
                max_path_length=cfg["horizon"], n_itr=n_itr, discount=cfg["discount"], gae_lambda=cfg["gae_lambda"],
                step_size=cfg["step_size"], optimizer=opt, sampler_args=dict(n_envs=cfg["lanes"], seed=seed))
    algo.start_worker()
    algo.init_opt()
    rets, bt, std = [], [], []
    for itr in range(n_itr):
        algo.train_itr(itr)
        tab = logger.get_last_table()
        rets.append(tab["AverageReturn"]), bt.append(algo.optimizer.last_info["n_iter"]), std.append(tab["AveragePolicyStd"])
    print("%s %s seed %d: last-5 mean %.3f  (itr 9: %.2f, 19: %.2f, 29: %.2f, 39: %.2f) backtracks %.2f policy std %.3f -> %.3f" %
          (env_name, variant, seed, np.mean(rets[-5:]), rets[9], rets[19], rets[29], rets[-1], np.mean(bt), std[0], std[-1]),
          flush=True)


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 7, int(sys.argv[3]) if len(sys.argv) > 3 else 40,
         sys.argv[4] if len(sys.argv) > 4 else "swimmer")

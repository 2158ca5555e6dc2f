"""Generate tests/golden/oracle_swimmer_trpo_curve.json: the float64 ORACLE's TRPO learning curve on Swimmer with the lane
semantics and the Philox noise streams of the CUDA path (same seed / iteration keys, so iteration 0 sees the same noise as
LaneSampler(seed=SEED) on the GPU).  Purpose: north_star's "TRPO AverageReturn on Swimmer within +-5 % of reference at
matched sample count" -- the reference's MuJoCo Swimmer cannot run in this container, so the comparison
target is the oracle restatement (oracle/planar.py), at N lanes x T steps per iteration.

Run (CPU only, ~2 min per iteration):  python tests/golden/make_swimmer_curve.py [n_itr] [swimmer|hopper] [sampler seed]
(hopper: the cfg4 net (64,64), same keys; written to oracle_hopper_trpo_curve.json)
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import envs as E, optim as OPT, philox as PH, policy as P, sampler as S  # noqa: E402

N, T, SEED, POLICY_SEED = 1024, 500, 7, 3
DISCOUNT, GAE_LAMBDA, STEP_SIZE, CG_ITERS = 0.99, 1.0, 0.01, 10


def main(n_itr, env_name="swimmer", seed=None):
    global SEED
    HIDDEN = 64 if env_name == "hopper" else 32
    OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_%s_trpo_curve.json" % env_name)
    if seed is not None and seed != SEED:       # extra sampler seeds (same initial policy): run-to-run spread of the curve
        SEED = seed
        OUT = OUT.replace(".json", "_seed%d.json" % seed)
    env = E.make(env_name)
    dims = P.Dims(env.O, (HIDDEN, HIDDEN), env.A)
    theta = P.init_params(dims, np.random.RandomState(POLICY_SEED))      # == GaussianMLPPolicy(..., seed=POLICY_SEED)
    coeffs = None
    rows = []
    for itr in range(n_itr):
        t0 = time.time()
        Ae, Ke = env.A + (env.A & 1), env.K + (env.K & 1)        # Box-Muller pairs: pad odd counts, drop the extra
        eps = PH.normal_from_raw(PH.raw_block(T, 0, Ae, N, 0, SEED, itr, 0))[:, :env.A]
        rr = PH.normal_from_raw(PH.raw_block(T + 1, 0, Ke, N, 0, SEED, itr, 1))[:, :env.K]
        traj = S.rollout_lanes(env, theta, dims, N, T, T, eps, rr)
        out = S.process_samples_lanes(traj, coeffs, DISCOUNT, GAE_LAMBDA, center_adv=True)
        coeffs = S.lfb_fit_lanes(traj["obs"], traj["tstep"], out["ret"])
        batch = S.batch_from_traj(traj, out["adv"])
        theta, info = OPT.trpo_step(theta, batch, dims, step_size=STEP_SIZE, cg_iters=CG_ITERS)
        st = out["stats"]
        rows.append(dict(itr=itr, AverageReturn=float(st["AverageReturn"]), StdReturn=float(st["StdReturn"]),
                         AverageDiscountedReturn=float(st["AverageDiscountedReturn"]), NumTrajs=int(st["NumTrajs"]),
                         MeanKL=float(info["constraint_val"]), LossBefore=float(info["loss_before"]),
                         LossAfter=float(info["loss"]), backtrack_iters=int(info["n_iter"]),
                         rejected=bool(info["rejected"]), seconds=time.time() - t0))
        print(rows[-1], flush=True)
        with open(OUT, "w") as f:
            json.dump(dict(config=dict(env=env_name, lanes=N, horizon=T, seed=SEED, policy_seed=POLICY_SEED,
                                       hidden=[HIDDEN, HIDDEN], discount=DISCOUNT, gae_lambda=GAE_LAMBDA,
                                       step_size=STEP_SIZE, cg_iters=CG_ITERS, baseline="LinearFeatureBaseline",
                                       arithmetic="float64 NumPy oracle (oracle/planar.py, oracle/optim.py)"),
                           curve=rows), f, indent=1)


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 40, sys.argv[2] if len(sys.argv) > 2 else "swimmer",
         int(sys.argv[3]) if len(sys.argv) > 3 else None)

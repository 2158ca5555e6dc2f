"""Generate tests/golden/reference_golden.npz by running the REAL reference code
(/root/reference, rll/rllab @ ba78e4c) in the build container.

Run:  python tests/golden/make_golden.py
The GPU box has no /root/reference: tests only read the committed .npz.

What is pinned (each key prefix = one reference entry point, run verbatim via oracle/ref_shims.py):
  ps_*    rllab.sampler.base.BaseSampler.process_samples  (+ LinearFeatureBaseline predict/fit)
  cg_*    rllab.misc.krylov.cg
  opt_*   rllab.optimizers.conjugate_gradient_optimizer.ConjugateGradientOptimizer.optimize
          (callables injected into _opt_fun / _hvp_approach; the MLP math inside them is the oracle's)
  dg_*    rllab.distributions.diagonal_gaussian.DiagonalGaussian.{kl,log_likelihood,entropy}
  pt_*    rllab.sampler.utils.rollout over normalize(PointEnv()) (examples/point_env.py)
  tr_*    rllab.sampler.parallel_sampler.truncate_paths
  misc_*  special.discount_cumsum, special.explained_variance_1d, algos.util.center_advantages
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True

from oracle import ref_shims  # noqa: E402
from oracle import policy as P  # noqa: E402
from oracle import sampler as S  # noqa: E402


def main():
    ref = ref_shims.import_reference()
    import rllab.misc.logger as logger
    out = {}
    rng = np.random.RandomState(1234)

    # ------------------------------------------------------------------ process_samples
    O, A, T, N = 3, 2, 23, 7
    obs = rng.randn(O, T, N) * 4.0      # some |o| > 10 after scaling below to exercise the clip
    obs[0] *= 4.0
    rew = rng.randn(T, N)
    flags = np.zeros((T, N), np.uint8)
    tstep = np.zeros((T, N), np.uint16)
    max_path_length = 9
    for n in range(N):
        plen = 0
        for t in range(T):
            tstep[t, n] = plen
            plen += 1
            done = rng.rand() < 0.12
            end = done or plen >= max_path_length or t == T - 1
            flags[t, n] = (1 if done else 0) + (2 if end else 0)
            if end:
                plen = 0
    traj = dict(obs=obs, act=rng.randn(A, T, N), mean=rng.randn(A, T, N), rew=rew, flags=flags,
                tstep=tstep, log_std=np.array([-0.3, 0.2]))
    coeffs_prev = rng.randn(2 * O + 4) * 0.1
    for tag, coeffs, lam, center, positive in (("a", None, 1.0, True, False),
                                              ("b", coeffs_prev, 0.97, True, False),
                                              ("c", coeffs_prev, 0.9, False, True)):
        paths = S.lanes_to_paths(traj)
        for p in paths:
            p.pop("_lane"), p.pop("_t0")
        baseline = ref.lfb.LinearFeatureBaseline(env_spec=None)
        if coeffs is not None:
            baseline.set_param_values(coeffs.copy())

        class _Pol(object):
            recurrent = False
            distribution = ref.diagonal_gaussian.DiagonalGaussian(A)

        class _Algo(object):
            pass
        algo = _Algo()
        algo.baseline, algo.policy = baseline, _Pol()
        algo.discount, algo.gae_lambda = 0.99, lam
        algo.center_adv, algo.positive_adv = center, positive
        tab = {}
        orig = logger.record_tabular
        logger.record_tabular = lambda k, v: tab.__setitem__(k, v)
        try:
            sd = ref.sampler_base.BaseSampler(algo).process_samples(0, paths)
        finally:
            logger.record_tabular = orig
        # scatter the reference's path-major outputs back to the lane layout
        adv = np.zeros((T, N))
        ret = np.zeros((T, N))
        k = 0
        for p in S.lanes_to_paths(traj):
            L = len(p["rewards"])
            adv[p["_t0"]:p["_t0"] + L, p["_lane"]] = sd["advantages"][k:k + L]
            ret[p["_t0"]:p["_t0"] + L, p["_lane"]] = sd["returns"][k:k + L]
            k += L
        out["ps_%s_adv" % tag] = adv
        out["ps_%s_ret" % tag] = ret
        out["ps_%s_fit" % tag] = np.asarray(baseline.get_param_values())
        for key in ("AverageDiscountedReturn", "AverageReturn", "ExplainedVariance", "NumTrajs", "Entropy",
                    "Perplexity", "StdReturn", "MaxReturn", "MinReturn"):
            out["ps_%s_%s" % (tag, key)] = np.float64(tab[key])
        out["ps_%s_cfg" % tag] = np.array([0.99, lam, float(center), float(positive)])
    for k_, v_ in traj.items():
        out["ps_in_" + k_] = v_
    out["ps_in_coeffs_prev"] = coeffs_prev

    # ------------------------------------------------------------------ krylov.cg
    n = 12
    M = rng.randn(n, n)
    Aspd = M @ M.T + 0.5 * np.eye(n)
    b = rng.randn(n)
    out["cg_A"], out["cg_b"] = Aspd, b
    out["cg_x10"] = ref.krylov.cg(lambda x: Aspd @ x, b.copy(), cg_iters=10)
    out["cg_x3"] = ref.krylov.cg(lambda x: Aspd @ x, b.copy(), cg_iters=3)

    # ------------------------------------------------------------------ ConjugateGradientOptimizer.optimize
    dims = P.Dims(3, (8, 8), 2)
    B = 257
    theta0 = P.init_params(dims, np.random.RandomState(7))
    batch = dict(obs=rng.randn(B, 3), adv=rng.randn(B))
    mu, ls = P.forward(theta0, batch["obs"], dims)
    batch["old_mean"], batch["old_log_std"] = mu, ls
    batch["actions"] = mu + np.exp(ls) * rng.randn(B, 2)
    for tag, step_size, scale in (("acc", 0.01, 1.0), ("small", 1e-4, 1.0), ("rej", 0.01, -1.0)):
        # scale=-1 flips the gradient handed to the optimizer -> ascent direction -> rejected step
        class _Target(object):
            def __init__(self):
                self.v = theta0.copy()

            def get_param_values(self, **tags):
                return self.v.copy()

            def set_param_values(self, v, **tags):
                self.v = np.array(v, dtype=np.float64)
        tgt = _Target()
        opt = ref.cg_opt.ConjugateGradientOptimizer(cg_iters=10, reg_coeff=1e-5)
        opt._target = tgt
        opt._max_constraint_val = step_size
        opt._constraint_name = "mean_kl"
        opt._opt_fun = dict(
            f_loss=lambda *a: P.surr_loss_trpo(tgt.v, batch, dims),
            f_grad=lambda *a: scale * P.grad_surr(tgt.v, batch, dims, "trpo"),
            f_loss_constraint=lambda *a: [P.surr_loss_trpo(tgt.v, batch, dims), P.kl_stats(tgt.v, batch, dims)[0]],
        )

        class _Hvp(object):
            def build_eval(self, inputs):
                return lambda x: P.fvp(theta0, batch, x, dims, 1e-5)
        opt._hvp_approach = _Hvp()
        orig_log = logger.log
        logger.log = lambda *a, **k: None
        try:
            opt.optimize([np.zeros((B, 1))])
        finally:
            logger.log = orig_log
        out["opt_%s_theta" % tag] = tgt.v.copy()
        out["opt_%s_cfg" % tag] = np.array([step_size, scale])
    out["opt_theta0"] = theta0
    for k_ in ("obs", "adv", "old_mean", "old_log_std", "actions"):
        out["opt_in_" + k_] = batch[k_]

    # ------------------------------------------------------------------ DiagonalGaussian numeric twins
    dg = ref.diagonal_gaussian.DiagonalGaussian(2)
    om, nm = rng.randn(31, 2), rng.randn(31, 2)
    ol, nl = rng.randn(31, 2) * 0.3, rng.randn(31, 2) * 0.3
    xs = rng.randn(31, 2)
    out["dg_om"], out["dg_nm"], out["dg_ol"], out["dg_nl"], out["dg_xs"] = om, nm, ol, nl, xs
    out["dg_kl"] = dg.kl(dict(mean=om, log_std=ol), dict(mean=nm, log_std=nl))
    out["dg_ll"] = dg.log_likelihood(xs, dict(mean=nm, log_std=nl))
    out["dg_ent"] = dg.entropy(dict(log_std=nl))

    # ------------------------------------------------------------------ rollout over normalize(PointEnv())
    Tp = 40
    acts = rng.randn(Tp, 2) * 0.8

    class _Replay(object):
        def __init__(self):
            self.t = 0

        def reset(self):
            self.t = 0

        def get_action(self, o):
            a = acts[self.t]
            self.t += 1
            return a, dict(mean=a * 0, log_std=a * 0)
    env = ref.normalized_env.normalize(ref.point_env.PointEnv())
    np.random.seed(42)
    s0 = np.random.uniform(-1, 1, size=(2,))
    np.random.seed(42)          # PointEnv.reset draws the same state again
    path = ref.sampler_utils.rollout(env, _Replay(), max_path_length=Tp)
    out["pt_f64_s0"] = s0
    out["pt_f64_actions"] = acts
    out["pt_f64_obs"] = np.asarray(path["observations"])
    out["pt_f64_rew"] = np.asarray(path["rewards"])
    assert np.array_equal(path["observations"][0], s0)
    # a path that terminates early (done): start next to the origin
    env = ref.normalized_env.normalize(ref.point_env.PointEnv())
    inner = env.wrapped_env
    inner.reset = lambda: (setattr(inner, "_state", np.array([0.05, -0.03])), np.copy(inner._state))[1]
    acts_d = np.tile(np.array([[-0.25, 0.15]]), (10, 1))

    class _Replay2(object):
        t = 0

        def reset(self):
            self.t = 0

        def get_action(self, o):
            a = acts_d[self.t]
            self.t += 1
            return a, dict()
    path = ref.sampler_utils.rollout(env, _Replay2(), max_path_length=10)
    out["pt_done_len"] = np.int64(len(path["rewards"]))
    out["pt_done_obs"] = path["observations"]
    out["pt_done_rew"] = path["rewards"]
    out["pt_done_actions"] = acts_d

    # ------------------------------------------------------------------ truncate_paths
    lens = [100, 50, 7, 33, 60]
    for ms in (130, 150, 1, 249, 250, 400):
        paths = [dict(observations=np.zeros((l, 1)), actions=np.zeros((l, 1)), rewards=np.zeros(l),
                      env_infos=dict(), agent_infos=dict(lala=np.zeros(l))) for l in lens]
        tp = ref.parallel_sampler.truncate_paths(paths, ms)
        out["tr_%d" % ms] = np.array([len(p["rewards"]) for p in tp], np.int64)
    out["tr_lens"] = np.array(lens, np.int64)

    # ------------------------------------------------------------------ misc
    x = rng.randn(50)
    out["misc_x"] = x
    out["misc_dcs"] = ref.special.discount_cumsum(x, 0.97)
    y = rng.randn(50)
    out["misc_y"] = y
    out["misc_ev"] = np.float64(ref.special.explained_variance_1d(x, y))
    out["misc_center"] = ref.algo_util.center_advantages(x)

    path_out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_golden.npz")
    np.savez_compressed(path_out, **out)
    print("wrote", path_out, "keys:", len(out), "bytes:", os.path.getsize(path_out))


if __name__ == "__main__":
    main()

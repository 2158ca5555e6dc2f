"""CPU: how well-posed is "parameters within 1e-5 after one TRPO step"?  Measured on the float64 oracle alone: the
relative change of the updated parameters when the Hessian-vector product is perturbed by relative noise of size eta
(eta = 1e-16 is a single float64 ulp -- e.g. a different BLAS summation order in the reference itself).

Result (CartPole, 1024 lanes x 50 steps, reg 1e-5): with cg_iters <= 6 the step is reproducible to 1e-12; with 8
iterations to ~1e-6; with the default 10 iterations a 1-ulp perturbation already moves the parameters by ~1e-5 and
1e-13 by ~4e-3.  Hence the GPU parity tests assert 1e-5 for cg_iters <= 4 (float32 kernels) / <= 8 (float64 parity
mode) and treat 10 iterations behaviourally (tests/test_gpu_algos.py, DESIGN.md "Parity limit")."""
import numpy as np

from oracle import envs as E, optim as OPT, policy as P, sampler as S


def _setup():
    env = E.make("cartpole")
    dims = P.Dims(4, (32, 32), 1)
    theta = P.init_params(dims, np.random.RandomState(3))
    N, T = 1024, 50
    rng = np.random.RandomState(0)
    traj = S.rollout_lanes(env, theta, dims, N, T, 50, rng.randn(T, 1, N), rng.rand(T + 1, 4, N))
    ps = S.process_samples_lanes(traj, None, 0.99, 1.0)
    return dims, theta, S.batch_from_traj(traj, ps["adv"])


def _step(dims, theta, batch, eta, cg_iters):
    r = np.random.RandomState(0)

    def f_Hx(th, x):
        h = P.fvp(th, batch, x, dims, 1e-5)
        return h * (1 + eta * r.randn(h.size))
    return OPT.trpo_optimize(lambda th: P.surr_loss_trpo(th, batch, dims), lambda th: P.grad_surr(th, batch, dims, "trpo"),
                             lambda th: (P.surr_loss_trpo(th, batch, dims), P.kl_stats(th, batch, dims)[0]), f_Hx, theta,
                             0.01, cg_iters)[0]


def test_trpo_step_sensitivity_to_hvp_rounding():
    dims, theta, batch = _setup()
    rel = lambda a, b: np.abs(a - b).max() / np.abs(a).max()
    base4, base10 = _step(dims, theta, batch, 0.0, 4), _step(dims, theta, batch, 0.0, 10)
    assert rel(base4, _step(dims, theta, batch, 1e-13, 4)) < 1e-9        # 4 iterations: well conditioned
    assert rel(base4, _step(dims, theta, batch, 3e-8, 4)) < 1e-5         # even at float32-level Hx error
    assert rel(base10, _step(dims, theta, batch, 1e-13, 10)) > 1e-5      # 10 iterations: 1e-13 noise breaks 1e-5
    assert rel(base10, _step(dims, theta, batch, 3e-8, 10)) > 1e-2       # float32-level Hx error: O(1) differences

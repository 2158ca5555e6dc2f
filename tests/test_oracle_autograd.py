"""Self-check of the oracle's stand-in for Theano autodiff (SURVEY 8c tier 2; VERDICT r01 "weak" item 3).

The reference builds the flat gradient with theano.grad (conjugate_gradient_optimizer.py:184-186,
first_order_optimizer.py:62-64) and the Hessian-vector product as grad(grad(mean_kl) . x)
(PerlmutterHvp, conjugate_gradient_optimizer.py:27-38) or as a central finite difference of the gradient
(FiniteDifferenceHvp, :77-97).  Theano is not installable here, so oracle/policy.py restates both by hand
(manual back-propagation; closed-form Gauss-Newton product).  These tests pin the restatement against two
independent implementations of the same definitions:
  * torch.autograd (float64, CPU; used as a CHECKER only -- nothing under rllab_b200/ imports autograd),
    with the double-backward written exactly as the reference's Theano graph, and
  * a central finite difference of the oracle's own loss / mean-KL gradient.
"""
import numpy as np
import pytest
import torch

from oracle import policy as P

LOG2PI = float(np.log(2.0 * np.pi))


def _problem(seed, O, H, A, B):
    rng = np.random.RandomState(seed)
    dims = P.Dims(O, H, A)
    theta = P.init_params(dims, rng) + 0.05 * rng.randn(dims.P)
    theta[-A:] = rng.uniform(-0.7, 0.3, size=A)                  # a non-trivial log_std
    obs = rng.randn(B, O)
    old_mean, old_log_std = P.forward(theta, obs, dims)
    actions = old_mean + np.exp(old_log_std) * rng.randn(B, A)
    batch = dict(obs=obs, actions=actions, adv=rng.randn(B), old_mean=old_mean, old_log_std=old_log_std)
    return dims, theta, batch, rng


def _t_forward(th, obs, dims, min_std=1e-6):
    k, ts = 0, []
    for s in dims.shapes:
        n = int(np.prod(s))
        ts.append(th[k:k + n].reshape(s))
        k += n
    h = obs
    nl = len(dims.H)
    for i in range(nl):
        h = torch.tanh(h @ ts[2 * i] + ts[2 * i + 1])
    mean = h @ ts[2 * nl] + ts[2 * nl + 1]
    log_std = torch.maximum(ts[-1], torch.tensor(np.log(min_std), dtype=th.dtype))
    return mean, log_std


def _t_logli(a, mean, log_std):
    z = (a - mean) / torch.exp(log_std)
    return -torch.sum(log_std * torch.ones_like(mean), -1) - 0.5 * torch.sum(z * z, -1) - 0.5 * mean.shape[-1] * LOG2PI


def _t_kl(om, ols, nm, nls):
    num = (om - nm) ** 2 + torch.exp(ols) ** 2 - torch.exp(nls) ** 2
    den = 2 * torch.exp(nls) ** 2 + 1e-8
    return torch.sum(num / den + nls - ols, -1)


def _t_batch(batch):
    return {k: torch.tensor(np.asarray(v), dtype=torch.float64) for k, v in batch.items()}


def _t_loss(th, tb, dims, kind):
    mean, log_std = _t_forward(th, tb["obs"], dims)
    lp = _t_logli(tb["actions"], mean, log_std)
    if kind == "trpo":                                            # npo.py:72-82
        lp_old = _t_logli(tb["actions"], tb["old_mean"], tb["old_log_std"])
        return -torch.mean(torch.exp(lp - lp_old) * tb["adv"])
    return -torch.mean(lp * tb["adv"])                            # vpg.py:91


def _t_mean_kl(th, tb, dims):
    mean, log_std = _t_forward(th, tb["obs"], dims)
    return torch.mean(_t_kl(tb["old_mean"], tb["old_log_std"] * torch.ones_like(mean), mean,
                            log_std * torch.ones_like(mean)))


CASES = [(0, 4, (8, 8), 2, 257), (1, 13, (32, 32), 2, 64), (2, 20, (64, 64), 3, 48)]


@pytest.mark.parametrize("seed,O,H,A,B", CASES)
@pytest.mark.parametrize("kind", ["trpo", "vpg"])
def test_grad_surr_matches_autograd(seed, O, H, A, B, kind):
    dims, theta, batch, rng = _problem(seed, O, H, A, B)
    if kind == "trpo":                        # evaluate away from theta_old so that the likelihood ratio is not 1
        theta = theta + 0.01 * rng.randn(dims.P)
    th = torch.tensor(theta, dtype=torch.float64, requires_grad=True)
    g_ref, = torch.autograd.grad(_t_loss(th, _t_batch(batch), dims, kind), th)
    g = P.grad_surr(theta, batch, dims, kind)
    np.testing.assert_allclose(g, g_ref.numpy(), rtol=1e-9, atol=1e-12 * np.abs(g_ref.numpy()).max() + 1e-15)
    loss = P.surr_loss_trpo(theta, batch, dims) if kind == "trpo" else P.surr_loss_vpg(theta, batch, dims)
    np.testing.assert_allclose(loss, float(_t_loss(th, _t_batch(batch), dims, kind).detach()), rtol=1e-12)


@pytest.mark.parametrize("seed,O,H,A,B", CASES)
def test_fvp_matches_double_backward_of_mean_kl(seed, O, H, A, B):
    """PerlmutterHvp: Hx = grad(sum(grad(mean_kl, params) * x), params) + reg * x at theta == theta_old."""
    dims, theta, batch, rng = _problem(seed, O, H, A, B)
    x = rng.randn(dims.P)
    tb = _t_batch(batch)
    th = torch.tensor(theta, dtype=torch.float64, requires_grad=True)
    g, = torch.autograd.grad(_t_mean_kl(th, tb, dims), th, create_graph=True)
    Hx_ref, = torch.autograd.grad(torch.sum(g * torch.tensor(x)), th)
    reg = 1e-5
    Hx = P.fvp(theta, batch, x, dims, reg)
    ref = Hx_ref.numpy() + reg * x
    np.testing.assert_allclose(Hx, ref, rtol=1e-9, atol=1e-12 * np.abs(ref).max())
    # the gradient of mean_kl vanishes at theta_old (the premise of the Gauss-Newton closed form): exactly for the mean
    # network, up to the 1e-8 of the KL denominator (diagonal_gaussian.py:28) for log_std: eps / (2 sigma^2 + eps)
    gk = g.detach().numpy()
    assert np.abs(gk[:-A]).max() < 1e-12 and np.abs(gk[-A:]).max() < 1e-7


@pytest.mark.parametrize("seed,O,H,A,B", CASES[:2])
def test_fvp_matches_finite_difference_of_kl_gradient(seed, O, H, A, B):
    """FiniteDifferenceHvp (symmetric form, conjugate_gradient_optimizer.py:77-97):
    Hx ~ (grad_kl(theta + eps x) - grad_kl(theta - eps x)) / (2 eps), eps = base_eps / ||x||."""
    dims, theta, batch, rng = _problem(seed, O, H, A, B)
    x = rng.randn(dims.P)
    tb = _t_batch(batch)

    def grad_kl(v):
        th = torch.tensor(v, dtype=torch.float64, requires_grad=True)
        g, = torch.autograd.grad(_t_mean_kl(th, tb, dims), th)
        return g.numpy()
    eps = 1e-5 / np.linalg.norm(x)
    fd = (grad_kl(theta + eps * x) - grad_kl(theta - eps * x)) / (2 * eps)
    Hx = P.fvp(theta, batch, x, dims, 0.0)
    np.testing.assert_allclose(Hx, fd, rtol=0, atol=2e-7 * np.abs(Hx).max())


@pytest.mark.parametrize("seed,O,H,A,B", CASES[:2])
@pytest.mark.parametrize("kind", ["trpo", "vpg"])
def test_grad_surr_matches_finite_difference_of_loss(seed, O, H, A, B, kind):
    dims, theta, batch, rng = _problem(seed, O, H, A, B)
    theta = theta + 0.01 * rng.randn(dims.P)
    f = (lambda v: P.surr_loss_trpo(v, batch, dims)) if kind == "trpo" else (lambda v: P.surr_loss_vpg(v, batch, dims))
    g = P.grad_surr(theta, batch, dims, kind)
    idx = rng.choice(dims.P, size=40, replace=False)
    h = 1e-6
    for i in idx:
        e = np.zeros(dims.P)
        e[i] = h
        fd = (f(theta + e) - f(theta - e)) / (2 * h)
        assert abs(fd - g[i]) <= 1e-7 * max(1.0, np.abs(g).max()), (i, fd, g[i])


def test_min_std_clamp_routes_gradient_like_maximum():
    """TT.maximum(param, log(min_std)) (gaussian_mlp_policy.py:100-101): zero gradient / zero FVP block on clamped
    log_std entries -- same convention as torch.maximum away from ties."""
    dims, theta, batch, rng = _problem(3, 4, (8, 8), 2, 33)
    theta[-2] = np.log(1e-6) - 1.0                       # clamped entry
    batch["old_mean"], batch["old_log_std"] = P.forward(theta, batch["obs"], dims)
    th = torch.tensor(theta, dtype=torch.float64, requires_grad=True)
    g_ref, = torch.autograd.grad(_t_loss(th, _t_batch(batch), dims, "vpg"), th)
    g = P.grad_surr(theta, batch, dims, "vpg")
    assert g[-2] == 0.0 and g_ref.numpy()[-2] == 0.0
    np.testing.assert_allclose(g, g_ref.numpy(), rtol=1e-8, atol=1e-12 * np.abs(g).max())

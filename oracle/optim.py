"""Oracle: conjugate gradient, TRPO step (CG + backtracking line search), VPG/Adam step.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates rllab/misc/krylov.py:7-39 and
rllab/optimizers/conjugate_gradient_optimizer.py:229-296; both are pinned against the
reference's own code in tests/golden (make_golden.py runs the real modules).
"""
import numpy as np

from . import policy as P


def cg(f_Ax, b, cg_iters=10, residual_tol=1e-10):
    """krylov.py:7-39 (Demmel p.312)."""
    p = b.copy()
    r = b.copy()
    x = np.zeros_like(b)
    rdotr = r.dot(r)
    for _ in range(cg_iters):
        z = f_Ax(p)
        v = rdotr / p.dot(z)
        x += v * p
        r -= v * z
        newrdotr = r.dot(r)
        mu = newrdotr / rdotr
        p = r + mu * p
        rdotr = newrdotr
        if rdotr < residual_tol:
            break
    return x


def trpo_optimize(f_loss, f_grad, f_loss_constraint, f_Hx, theta, max_constraint_val=0.01,
                  cg_iters=10, backtrack_ratio=0.8, max_backtracks=15, accept_violation=False):
    """conjugate_gradient_optimizer.py:229-296 with callables of theta (f_Hx(theta, x) already
    includes reg_coeff * x, as PerlmutterHvp.build_eval does at :48-55).
    Returns (theta_new, info)."""
    theta = np.asarray(theta, dtype=np.float64)
    loss_before = f_loss(theta)
    flat_g = f_grad(theta)
    Hx = lambda x: f_Hx(theta, x)
    descent_direction = cg(Hx, flat_g, cg_iters=cg_iters)
    with np.errstate(invalid="ignore", divide="ignore"):
        initial_step_size = np.sqrt(
            2.0 * max_constraint_val * (1. / (descent_direction.dot(Hx(descent_direction)) + 1e-8)))
    if np.isnan(initial_step_size):
        initial_step_size = 1.
    flat_descent_step = initial_step_size * descent_direction
    prev_param = np.copy(theta)
    n_iter = 0
    loss = constraint_val = np.nan
    cur_param = prev_param
    for n_iter, ratio in enumerate(backtrack_ratio ** np.arange(max_backtracks)):
        cur_step = ratio * flat_descent_step
        cur_param = prev_param - cur_step
        loss, constraint_val = f_loss_constraint(cur_param)
        if loss < loss_before and constraint_val <= max_constraint_val:
            break
    rejected = False
    if (np.isnan(loss) or np.isnan(constraint_val) or loss >= loss_before or
            constraint_val >= max_constraint_val) and not accept_violation:
        cur_param = prev_param
        rejected = True
    info = dict(loss_before=loss_before, loss=loss, constraint_val=constraint_val, n_iter=n_iter,
                rejected=rejected, initial_step_size=initial_step_size, flat_g=flat_g,
                descent_direction=descent_direction)
    return cur_param, info


def trpo_step(theta, batch, dims, step_size=0.01, cg_iters=10, reg_coeff=1e-5, backtrack_ratio=0.8,
              max_backtracks=15, min_std=1e-6, accept_violation=False):
    """One NPO/TRPO policy update on `batch` (npo.py:102-123 minus logging)."""
    f_loss = lambda th: P.surr_loss_trpo(th, batch, dims, min_std)
    f_grad = lambda th: P.grad_surr(th, batch, dims, "trpo", min_std)
    f_lc = lambda th: (P.surr_loss_trpo(th, batch, dims, min_std), P.kl_stats(th, batch, dims, min_std)[0])
    f_Hx = lambda th, x: P.fvp(th, batch, x, dims, reg_coeff, min_std)
    return trpo_optimize(f_loss, f_grad, f_lc, f_Hx, theta, step_size, cg_iters, backtrack_ratio,
                         max_backtracks, accept_violation)


def vpg_step(theta, batch, dims, adam_state, lr=1e-3, min_std=1e-6):
    """One VPG update: a single full-batch Adam step (vpg.py:110-130 with the defaults
    batch_size=None, max_epochs=1 of vpg.py:26-29; first_order_optimizer.py:84-133)."""
    g = P.grad_surr(theta, batch, dims, "vpg", min_std)
    m, v, t = adam_state
    theta, m, v, t = P.adam_step(np.asarray(theta, np.float64), g, m, v, t, lr=lr)
    return theta, (m, v, t)

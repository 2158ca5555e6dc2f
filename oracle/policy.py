"""Oracle: GaussianMLPPolicy + DiagonalGaussian + NPO/VPG surrogates (float64 NumPy).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates, with manual back-propagation instead of Theano autodiff:
  * rllab/policies/gaussian_mlp_policy.py:61-137  (mean MLP, ParamLayer log_std, min_std clamp)
  * rllab/core/network.py:36-81                   (MLP: tanh hidden, linear output, GlorotUniform/0 init)
  * rllab/core/lasagne_layers.py:9-30             (ParamLayer broadcast)
  * rllab/distributions/diagonal_gaussian.py:14-95
  * rllab/algos/npo.py:72-98, rllab/algos/vpg.py:80-107 (surrogates)
  * rllab/optimizers/conjugate_gradient_optimizer.py:22-55 (PerlmutterHvp == Gauss-Newton product at theta_old)

Flat parameter layout (core/lasagne_powered.py:16-20 + misc/tensor_utils.py:6-16, [3P ordering]):
  [W0 (O,h1) row-major, b0 (h1), W1 (h1,h2), b1 (h2), Wout (h2,A), bout (A), log_std (A)]
"""
import numpy as np

LOG2PI = np.log(2.0 * np.pi)


class Dims(object):
    def __init__(self, obs_dim, hidden_sizes, act_dim):
        self.O = int(obs_dim)
        self.H = tuple(int(h) for h in hidden_sizes)
        self.A = int(act_dim)
        sizes = (self.O,) + self.H + (self.A,)
        self.shapes = []
        for i in range(len(sizes) - 1):
            self.shapes.append((sizes[i], sizes[i + 1]))
            self.shapes.append((sizes[i + 1],))
        self.shapes.append((self.A,))
        self.P = int(sum(int(np.prod(s)) for s in self.shapes))


def unpack(flat, dims):
    out, k = [], 0
    for s in dims.shapes:
        n = int(np.prod(s))
        out.append(np.asarray(flat[k:k + n]).reshape(s))
        k += n
    assert k == dims.P
    return out


def pack(tensors):
    return np.concatenate([np.asarray(t).reshape(-1) for t in tensors])


def init_params(dims, rng, init_std=1.0):
    """GlorotUniform weights, zero biases, log_std = log(init_std)
    (core/network.py:38-39, gaussian_mlp_policy.py:88-94; Lasagne GlorotUniform [3P]:
    U(-a, a), a = sqrt(6 / (fan_in + fan_out)))."""
    ts = []
    for s in dims.shapes[:-1]:
        if len(s) == 2:
            a = np.sqrt(6.0 / (s[0] + s[1]))
            ts.append(rng.uniform(-a, a, size=s))
        else:
            ts.append(np.zeros(s))
    ts.append(np.full((dims.A,), np.log(init_std)))
    return pack(ts)


def forward(flat, obs, dims, min_std=1e-6, keep=False):
    """mean (n,A), log_std (A,) after the min_std clamp (gaussian_mlp_policy.py:98-101,118-122)."""
    ts = unpack(flat, dims)
    h = np.asarray(obs, dtype=np.float64)
    acts = [h]
    nl = len(dims.H)
    for i in range(nl):
        h = np.tanh(h @ ts[2 * i] + ts[2 * i + 1])
        acts.append(h)
    mean = h @ ts[2 * nl] + ts[2 * nl + 1]
    log_std_param = ts[-1]
    if min_std is not None:
        log_std = np.maximum(log_std_param, np.log(min_std))
    else:
        log_std = log_std_param
    if keep:
        return mean, log_std, acts
    return mean, log_std


def log_likelihood(actions, mean, log_std):
    """diagonal_gaussian.py:77-83"""
    zs = (actions - mean) / np.exp(log_std)
    return -np.sum(log_std * np.ones_like(mean), axis=-1) - 0.5 * np.sum(np.square(zs), axis=-1) \
        - 0.5 * mean.shape[-1] * LOG2PI


def kl(old_mean, old_log_std, new_mean, new_log_std):
    """diagonal_gaussian.py:36-56 (note the +1e-8 in the denominator)."""
    old_std = np.exp(old_log_std)
    new_std = np.exp(new_log_std)
    numerator = np.square(old_mean - new_mean) + np.square(old_std) - np.square(new_std)
    denominator = 2 * np.square(new_std) + 1e-8
    return np.sum(numerator / denominator + new_log_std - old_log_std, axis=-1)


def entropy(log_std):
    """diagonal_gaussian.py:85-87"""
    return np.sum(log_std + np.log(np.sqrt(2 * np.pi * np.e)), axis=-1)


# --------------------------------------------------------------------------------------
# losses.  batch = dict(obs (B,O), actions (B,A), adv (B,), old_mean (B,A), old_log_std (A,) or (B,A))
# --------------------------------------------------------------------------------------
def surr_loss_trpo(flat, batch, dims, min_std=1e-6):
    """npo.py:72-82: -mean(exp(logp_new - logp_old) * adv)"""
    mean, log_std = forward(flat, batch["obs"], dims, min_std)
    lp_new = log_likelihood(batch["actions"], mean, log_std)
    lp_old = log_likelihood(batch["actions"], batch["old_mean"], batch["old_log_std"])
    return -np.mean(np.exp(lp_new - lp_old) * batch["adv"])


def surr_loss_vpg(flat, batch, dims, min_std=1e-6):
    """vpg.py:91: -mean(logp * adv)"""
    mean, log_std = forward(flat, batch["obs"], dims, min_std)
    return -np.mean(log_likelihood(batch["actions"], mean, log_std) * batch["adv"])


def kl_stats(flat, batch, dims, min_std=1e-6):
    """(mean_kl, max_kl) of KL(old || new)   npo.py:73,80 ; vpg.py:92,98-99"""
    mean, log_std = forward(flat, batch["obs"], dims, min_std)
    k = kl(batch["old_mean"], batch["old_log_std"], mean, log_std * np.ones_like(mean))
    return np.mean(k), np.max(k)


def _backward(flat, dims, acts, dmean, dlog_std_param):
    """Back-propagate d(loss)/d(mean) (n,A) through the tanh MLP; returns the flat gradient."""
    ts = unpack(flat, dims)
    nl = len(dims.H)
    grads = [None] * len(ts)
    delta = dmean
    grads[2 * nl] = acts[nl].T @ delta
    grads[2 * nl + 1] = delta.sum(axis=0)
    for i in range(nl - 1, -1, -1):
        delta = (delta @ ts[2 * (i + 1)].T) * (1.0 - np.square(acts[i + 1]))
        grads[2 * i] = acts[i].T @ delta
        grads[2 * i + 1] = delta.sum(axis=0)
    grads[-1] = dlog_std_param
    return pack(grads)


def grad_surr(flat, batch, dims, kind, min_std=1e-6):
    """Flat gradient of the TRPO ('trpo') or VPG ('vpg') surrogate.  Stands in for
    theano.grad(loss, params) (conjugate_gradient_optimizer.py:184-186, first_order_optimizer.py:62-64)."""
    mean, log_std, acts = forward(flat, batch["obs"], dims, min_std, keep=True)
    B = mean.shape[0]
    std = np.exp(log_std)
    z = (batch["actions"] - mean) / std
    if kind == "trpo":
        lp_new = log_likelihood(batch["actions"], mean, log_std)
        lp_old = log_likelihood(batch["actions"], batch["old_mean"], batch["old_log_std"])
        w = np.exp(lp_new - lp_old) * batch["adv"]
    elif kind == "vpg":
        w = batch["adv"]
    else:
        raise ValueError(kind)
    coef = (-w / B)[:, None]
    dmean = coef * z / std
    dlog_std = (coef * (np.square(z) - 1.0)).sum(axis=0)
    if min_std is not None:
        # TT.maximum passes the gradient to the larger argument (param > log(min_std))
        ts = unpack(flat, dims)
        dlog_std = np.where(ts[-1] > np.log(min_std), dlog_std, 0.0)
    return _backward(flat, dims, acts, dmean, dlog_std)


def fvp(flat, batch, x, dims, reg_coeff=1e-5, min_std=1e-6):
    """Hx = grad(grad(mean_kl) . x) + reg*x at theta == theta_old, in closed form
    (PerlmutterHvp, conjugate_gradient_optimizer.py:22-55):
    J_mu^T diag(M_mu) J_mu x / B  (+)  diag(M_l) x_l ,  M_mu = 2/(2 s + 1e-8), M_l = 4 s (2 s - e)/(2 s + e)^2, s = sigma^2.
    Valid because the batch's old_mean/old_log_std were produced by `flat` itself."""
    ts = unpack(flat, dims)
    xs = unpack(x, dims)
    nl = len(dims.H)
    mean, log_std, acts = forward(flat, batch["obs"], dims, min_std, keep=True)
    B = mean.shape[0]
    # tangent forward
    dh = np.zeros_like(acts[0])
    for i in range(nl):
        pre = dh @ ts[2 * i] + acts[i] @ xs[2 * i] + xs[2 * i + 1]
        dh = (1.0 - np.square(acts[i + 1])) * pre
    dmu = dh @ ts[2 * nl] + acts[nl] @ xs[2 * nl] + xs[2 * nl + 1]
    s = np.exp(2.0 * log_std)
    eps = 1e-8
    M_mu = 2.0 / (2.0 * s + eps)
    M_l = 4.0 * s * (2.0 * s - eps) / np.square(2.0 * s + eps)
    dmean = dmu * M_mu / B
    dls = M_l * xs[-1]
    if min_std is not None:
        dls = np.where(ts[-1] > np.log(min_std), dls, 0.0)
    return _backward(flat, dims, acts, dmean, dls) + reg_coeff * np.asarray(x)


def mean_kl_at(flat_new, flat_old, obs, dims, min_std=1e-6):
    mo, lo = forward(flat_old, obs, dims, min_std)
    mn, ln = forward(flat_new, obs, dims, min_std)
    return np.mean(kl(mo, lo * np.ones_like(mo), mn, ln * np.ones_like(mn)))


def adam_step(flat, g, m, v, t, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8):
    """lasagne.updates.adam [3P Lasagne @484866c]: t<-t+1; a=lr*sqrt(1-b2^t)/(1-b1^t);
    m<-b1 m+(1-b1) g; v<-b2 v+(1-b2) g^2; theta<-theta - a m/(sqrt(v)+eps)
    (first_order_optimizer.py:21-22,43,62-65)."""
    t = t + 1
    a_t = lr * np.sqrt(1.0 - b2 ** t) / (1.0 - b1 ** t)
    m = b1 * m + (1.0 - b1) * g
    v = b2 * v + (1.0 - b2) * np.square(g)
    flat = flat - a_t * m / (np.sqrt(v) + eps)
    return flat, m, v, t

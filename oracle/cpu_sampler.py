"""Oracle: the reference's CPU sampler structure, restated -- per-path Python `rollout()` loops in a pool of worker
processes, then per-path process_samples and a full-batch NumPy policy update.  TEST / BASELINE INFRASTRUCTURE ONLY:
used by bench.py's `cpu_baseline` leg and `--impl reference` arm (the reference itself is Python + Theano/Box2D/MuJoCo
and cannot run on the GPU box).

Restates:
  * rllab/sampler/utils.py:6-43                 rollout(env, agent, max_path_length)
  * rllab/sampler/parallel_sampler.py:92-126    _worker_collect_one_path / sample_paths (collect until >= max_samples)
  * rllab/sampler/stateful_pool.py:102-157      run_collect (threshold semantics, whole paths)
  * rllab/policies/gaussian_mlp_policy.py:125-130  get_action (one forward per step, np.random.normal)
  * rllab/sampler/base.py:48-182                process_samples (per-path loops)
  * rllab/algos/vpg.py:110-130 / npo.py:102-123  update
"""
import multiprocessing as mp
import time

import numpy as np

from . import envs as E
from . import optim as OPT
from . import policy as P
from . import sampler as S

_G = {}


class _ScalarEnv(object):
    """Single-environment Env API over the lane restatement (n = 1), with NormalizedEnv.step semantics."""

    def __init__(self, name):
        self.e = E.make(name)
        self.s = None

    def reset(self):
        raw = np.random.uniform(size=(self.e.K, 1)) if self.e.noise_kind == "uniform" else \
            np.random.normal(size=(self.e.K, 1))
        self.s = self.e.reset(raw)
        return self.e.obs(self.s)[:, 0]

    def step(self, action):
        u = self.e.scale_action(np.asarray(action, np.float64).reshape(-1, 1))
        self.s, r, d = self.e.step(self.s, u)
        return self.e.obs(self.s)[:, 0], float(r[0]), bool(d[0]), {}


class _Policy(object):
    def __init__(self, dims, theta, min_std=1e-6):
        self.dims, self.min_std = dims, min_std
        self.set(theta)

    def set(self, theta):
        self.theta = np.asarray(theta, np.float64)
        self.ts = P.unpack(self.theta, self.dims)
        self.log_std = np.maximum(self.ts[-1], np.log(self.min_std))

    def get_action(self, o):
        h = o
        nl = len(self.dims.H)
        for i in range(nl):
            h = np.tanh(h @ self.ts[2 * i] + self.ts[2 * i + 1])
        mean = h @ self.ts[2 * nl] + self.ts[2 * nl + 1]
        rnd = np.random.normal(size=mean.shape)
        return rnd * np.exp(self.log_std) + mean, dict(mean=mean, log_std=self.log_std)


def rollout(env, agent, max_path_length):
    """sampler/utils.py:6-43"""
    observations, actions, rewards, means = [], [], [], []
    o = env.reset()
    path_length = 0
    while path_length < max_path_length:
        a, info = agent.get_action(o)
        next_o, r, d, _ = env.step(a)
        observations.append(o)
        rewards.append(r)
        actions.append(a)
        means.append(info["mean"])
        path_length += 1
        if d:
            break
        o = next_o
    L = len(rewards)
    return dict(observations=np.array(observations), actions=np.array(actions), rewards=np.array(rewards),
                agent_infos=dict(mean=np.array(means), log_std=np.tile(agent.log_std, (L, 1))), env_infos=dict())


def _worker_init(env_name, dims_args, seed):
    ident = mp.current_process()._identity
    np.random.seed(seed + (ident[0] if ident else 0))            # parallel_sampler.set_seed: seed + worker id
    _G["env"] = _ScalarEnv(env_name)
    _G["dims"] = P.Dims(*dims_args)
    _G["policy"] = None


def _worker_collect(args):
    theta, n_samples, max_path_length = args
    if _G["policy"] is None:
        _G["policy"] = _Policy(_G["dims"], theta)
    else:
        _G["policy"].set(theta)                                   # _worker_set_policy_params
    paths, got = [], 0
    while got < n_samples:
        p = rollout(_G["env"], _G["policy"], max_path_length)
        paths.append(p)
        got += len(p["rewards"])
    return paths


class CpuSampler(object):
    """n_parallel worker processes, each collecting whole paths until its share of max_samples is reached."""

    def __init__(self, env_name, dims, n_parallel, seed=1):
        self.env_name, self.dims, self.n_parallel = env_name, dims, n_parallel
        dims_args = (dims.O, dims.H, dims.A)
        if n_parallel > 1:
            self.pool = mp.get_context("fork").Pool(n_parallel, initializer=_worker_init,
                                                    initargs=(env_name, dims_args, seed))
        else:
            self.pool = None
            _worker_init(env_name, dims_args, seed)

    def sample_paths(self, theta, max_samples, max_path_length):
        if self.pool is None:
            return _worker_collect((theta, max_samples, max_path_length))
        share = -(-max_samples // self.n_parallel)
        res = self.pool.map(_worker_collect, [(theta, share, max_path_length)] * self.n_parallel)
        return sum(res, [])

    def close(self):
        if self.pool is not None:
            self.pool.terminate()
            self.pool.join()


def process_paths(paths, coeffs, discount, gae_lambda):
    """sampler/base.py:48-104 + linear_feature_baseline fit (base.py:163-167), per-path loops as in the reference."""
    for path in paths:
        b = S.lfb_features(path["observations"]).dot(coeffs) if coeffs is not None else np.zeros(len(path["rewards"]))
        pb = np.append(b, 0)
        deltas = path["rewards"] + discount * pb[1:] - pb[:-1]
        path["advantages"] = S.discount_cumsum(deltas, discount * gae_lambda)
        path["returns"] = S.discount_cumsum(path["rewards"], discount)
    obs = np.concatenate([p["observations"] for p in paths])
    act = np.concatenate([p["actions"] for p in paths])
    adv = np.concatenate([p["advantages"] for p in paths])
    mean = np.concatenate([p["agent_infos"]["mean"] for p in paths])
    adv = (adv - np.mean(adv)) / (adv.std() + 1e-8)
    featmat = np.concatenate([S.lfb_features(p["observations"]) for p in paths])
    returns = np.concatenate([p["returns"] for p in paths])
    new_coeffs = S.lfb_fit_normal(featmat.T.dot(featmat), featmat.T.dot(returns))
    batch = dict(obs=obs, actions=act, adv=adv, old_mean=mean, old_log_std=paths[0]["agent_infos"]["log_std"][0])
    avg_ret = float(np.mean([p["rewards"].sum() for p in paths]))
    return batch, new_coeffs, avg_ret


def run_iteration(sampler, theta, coeffs, dims, algo, max_samples, max_path_length, adam_state=None, discount=0.99,
                  gae_lambda=1.0):
    """One BatchPolopt iteration on the CPU.  Returns (theta, coeffs, adam_state, n_samples, seconds, avg_return)."""
    t0 = time.time()
    paths = sampler.sample_paths(theta, max_samples, max_path_length)
    batch, coeffs, avg_ret = process_paths(paths, coeffs, discount, gae_lambda)
    if algo == "vpg":
        if adam_state is None:
            adam_state = (np.zeros(dims.P), np.zeros(dims.P), 0)
        P.surr_loss_vpg(theta, batch, dims)                                   # loss_before (vpg.py:122)
        theta, adam_state = OPT.vpg_step(theta, batch, dims, adam_state)
        P.surr_loss_vpg(theta, batch, dims)                                   # loss_after
        P.kl_stats(theta, batch, dims)                                        # f_kl
    else:
        theta, _ = OPT.trpo_step(theta, batch, dims)
    return theta, coeffs, adam_state, len(batch["adv"]), time.time() - t0, avg_ret

"""NormalizedEnv / normalize (rllab/envs/normalized_env.py:11-103).  The action affine map + clip
(normalized_env.py:78-92) is executed inside the CUDA kernels (scale_action in csrc/envs.cuh); running
observation / reward normalisation (off by default in the reference, :16-17) is not part of the hot path and is
rejected loudly."""
import numpy as np

from ..spaces import Box
from .base import Env, Step


class ProxyEnv(Env):
    """rllab/envs/proxy_env.py"""

    def __init__(self, wrapped_env):
        self._wrapped_env = wrapped_env

    @property
    def wrapped_env(self):
        return self._wrapped_env

    def reset(self, **kwargs):
        return self._wrapped_env.reset(**kwargs)

    @property
    def action_space(self):
        return self._wrapped_env.action_space

    @property
    def observation_space(self):
        return self._wrapped_env.observation_space

    def step(self, action):
        return self._wrapped_env.step(action)

    def render(self, *args, **kwargs):
        return self._wrapped_env.render(*args, **kwargs)

    def log_diagnostics(self, paths, *args, **kwargs):
        self._wrapped_env.log_diagnostics(paths, *args, **kwargs)

    @property
    def horizon(self):
        return self._wrapped_env.horizon

    def terminate(self):
        self._wrapped_env.terminate()

    def get_param_values(self):
        return self._wrapped_env.get_param_values()

    def set_param_values(self, params):
        self._wrapped_env.set_param_values(params)


class NormalizedEnv(ProxyEnv):
    def __init__(self, env, scale_reward=1., normalize_obs=False, normalize_reward=False, obs_alpha=0.001,
                 reward_alpha=0.001):
        if normalize_obs or normalize_reward:
            raise NotImplementedError("running obs/reward normalisation is outside the B200 hot path "
                                      "(reference default is off: normalized_env.py:16-17)")
        if scale_reward != 1.:
            raise NotImplementedError("scale_reward != 1 is not supported by the fused kernels")
        ProxyEnv.__init__(self, env)
        self._scale_reward = scale_reward

    @property
    def env_kind(self):
        return self._wrapped_env.env_kind

    @property
    def action_space(self):
        if isinstance(self._wrapped_env.action_space, Box):
            ub = np.ones(self._wrapped_env.action_space.shape)
            return Box(-1 * ub, ub)
        return self._wrapped_env.action_space

    def step(self, action):
        inner = self._wrapped_env
        inner._normalized = True       # the kernel applies clip(lb + (a+1)/2 (ub-lb), lb, ub)
        try:
            next_obs, reward, done, info = inner.step(action)
        finally:
            inner._normalized = False
        return Step(next_obs, reward * self._scale_reward, done, **info)

    vectorized = True

    def vec_env_executor(self, n_envs, max_path_length):
        from .lane_env import VecEnvExecutor
        return VecEnvExecutor(self, n_envs, max_path_length)

    def __str__(self):
        return "Normalized: %s" % self._wrapped_env


normalize = NormalizedEnv

"""Env / Step / EnvSpec protocol (rllab/envs/base.py:6-100, rllab/envs/env_spec.py)."""
import collections


class EnvSpec(object):
    def __init__(self, observation_space, action_space):
        self._observation_space = observation_space
        self._action_space = action_space

    @property
    def observation_space(self):
        return self._observation_space

    @property
    def action_space(self):
        return self._action_space


class Env(object):
    def step(self, action):
        raise NotImplementedError

    def reset(self):
        raise NotImplementedError

    @property
    def action_space(self):
        raise NotImplementedError

    @property
    def observation_space(self):
        raise NotImplementedError

    @property
    def action_dim(self):
        return self.action_space.flat_dim

    def render(self):
        pass

    def log_diagnostics(self, paths):
        pass

    @property
    def spec(self):
        return EnvSpec(observation_space=self.observation_space, action_space=self.action_space)

    @property
    def horizon(self):
        raise NotImplementedError

    def terminate(self):
        pass

    def get_param_values(self):
        return None

    def set_param_values(self, params):
        pass


_Step = collections.namedtuple("Step", ["observation", "reward", "done", "info"])


def Step(observation, reward, done, **kwargs):
    return _Step(observation, reward, done, kwargs)

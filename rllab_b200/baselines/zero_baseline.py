"""ZeroBaseline (rllab/baselines/zero_baseline.py)."""
import numpy as np


class ZeroBaseline(object):
    def __init__(self, env_spec=None):
        pass

    def get_param_values(self, **kwargs):
        return None

    def set_param_values(self, val, **kwargs):
        pass

    def fit(self, paths):
        pass

    def predict(self, path):
        return np.zeros_like(path["rewards"])

    def log_diagnostics(self, paths):
        pass

    # device hooks used by the lane sampler
    def device_weights(self, obs_dim, device):
        return None

    def fit_lanes(self, batch, comm=None):
        pass

"""DiagonalGaussian, host-side numeric API of rllab/distributions/diagonal_gaussian.py:36-56,71-95
(kl / log_likelihood / sample / entropy / dist_info_keys).  The *_sym graph builders have no counterpart: their
compiled functions are the CUDA kernels of csrc/update.cu (b200rl_loss_kl / b200rl_grad / b200rl_fvp)."""
import numpy as np


class DiagonalGaussian(object):
    def __init__(self, dim):
        self._dim = dim

    @property
    def dim(self):
        return self._dim

    def kl(self, old_dist_info, new_dist_info):
        old_means, old_log_stds = old_dist_info["mean"], old_dist_info["log_std"]
        new_means, new_log_stds = new_dist_info["mean"], new_dist_info["log_std"]
        old_std, new_std = np.exp(old_log_stds), np.exp(new_log_stds)
        numerator = np.square(old_means - new_means) + np.square(old_std) - np.square(new_std)
        denominator = 2 * np.square(new_std) + 1e-8
        return np.sum(numerator / denominator + new_log_stds - old_log_stds, axis=-1)

    def log_likelihood(self, xs, dist_info):
        means, log_stds = dist_info["mean"], dist_info["log_std"]
        zs = (xs - means) / np.exp(log_stds)
        return - np.sum(log_stds, axis=-1) - 0.5 * np.sum(np.square(zs), axis=-1) - \
            0.5 * means.shape[-1] * np.log(2 * np.pi)

    def sample(self, dist_info):
        means, log_stds = dist_info["mean"], dist_info["log_std"]
        rnd = np.random.normal(size=means.shape)
        return rnd * np.exp(log_stds) + means

    def entropy(self, dist_info):
        log_stds = dist_info["log_std"]
        return np.sum(log_stds + np.log(np.sqrt(2 * np.pi * np.e)), axis=-1)

    @property
    def dist_info_keys(self):
        return ["mean", "log_std"]

"""LinearFeatureBaseline (rllab/baselines/linear_feature_baseline.py:6-43).

Host API (fit(paths) / predict(path) on the reference's path dicts) is kept verbatim for callers that hold host
paths.  The hot path uses the device hooks: `device_weights` feeds b200rl_process_samples (predict), and `fit_lanes`
reduces the normal equations A^T A | A^T y with b200rl_lfb_gram (one all-reduce across GPUs) and solves the
regularised d x d system (d = 2*obs_dim+4 <= 44) on the device with b200rl_lfb_solve, including the reference's
10x-regularisation retry loop (:26-37) -- no host round trip; the coefficients are copied to the host only when
get_param_values() / predict(path) / pickling asks for them."""
import numpy as np


class LinearFeatureBaseline(object):
    def __init__(self, env_spec=None, reg_coeff=1e-5):
        self._coeffs = None
        self._reg_coeff = reg_coeff
        self._dev_w = None

    def set_param_values(self, val, **tags):
        self._coeffs = val
        self._dev_w = None

    def _features(self, path):
        o = np.clip(path["observations"], -10, 10)
        l = len(path["rewards"])
        al = np.arange(l).reshape(-1, 1) / 100.0
        return np.concatenate([o, o ** 2, al, al ** 2, al ** 3, np.ones((l, 1))], axis=1)

    def _solve(self, AtA, Aty):
        reg_coeff = self._reg_coeff
        for _ in range(5):
            self._coeffs = np.linalg.lstsq(AtA + reg_coeff * np.identity(AtA.shape[0]), Aty, rcond=None)[0]
            if not np.any(np.isnan(self._coeffs)):
                break
            reg_coeff *= 10
        self._dev_w = None

    def fit(self, paths):
        featmat = np.concatenate([self._features(path) for path in paths])
        returns = np.concatenate([path["returns"] for path in paths])
        self._solve(featmat.T.dot(featmat), featmat.T.dot(returns))

    def predict(self, path):
        coeffs = self.get_param_values()
        if coeffs is None:
            return np.zeros(len(path["rewards"]))
        return self._features(path).dot(coeffs)

    def log_diagnostics(self, paths):
        pass

    # ---- device hooks (no host round trip: Gram reduction, all-reduce and the d x d Cholesky solve stay on the GPU)
    def get_param_values(self, **tags):
        if self._coeffs is None and self._dev_w is not None:
            self._coeffs = self._dev_w.cpu().numpy()
        return self._coeffs

    def device_weights(self, obs_dim, device):
        if self._dev_w is None and self._coeffs is None:
            return None
        if self._dev_w is None or self._dev_w.device != device:
            import torch
            assert len(self._coeffs) == 2 * obs_dim + 4
            self._dev_w = torch.as_tensor(np.asarray(self._coeffs, dtype=np.float64)).to(device)
        return self._dev_w

    def gram_lanes(self, batch):
        """Normal equations A^T A | A^T y of this rank's (unmasked) samples into batch.gram (all-reduced by the sampler
        together with the advantage statistics)."""
        from .. import ops
        ops.lfb_gram(batch, batch.gram)

    def solve_lanes(self, batch):
        """d x d regularised solve on the device from the all-reduced batch.gram (linear_feature_baseline.py:26-37)."""
        import torch
        from .. import ops
        d = 2 * batch.O + 4
        if getattr(self, "_w_bufs", None) is None or self._w_bufs[0].device != batch.device or self._w_bufs[0].numel() != d:
            # two weight buffers used alternately: process_samples of the next iteration reads the one solved now
            self._w_bufs = [torch.zeros((d,), dtype=torch.float64, device=batch.device) for _ in range(2)]
            self._w_cur = 0
            self._solve_info = torch.zeros((3,), dtype=torch.float64, device=batch.device)
        self._w_cur ^= 1
        w = self._w_bufs[self._w_cur]
        ops.lfb_solve(batch.O, batch.gram, self._reg_coeff, w, self._solve_info)
        self._dev_w = w
        self._coeffs = None            # fetched lazily by get_param_values()

    def fit_lanes(self, batch, comm=None):
        self.gram_lanes(batch)
        if comm is not None:
            comm.all_reduce_sum(batch.gram)
        self.solve_lanes(batch)

    def last_fit_info(self):
        """(final regularisation, attempts used, ok flag) of the last device solve; ok == 0 means all 5 attempts failed
        and the weights were set to zero (no baseline) rather than left undefined."""
        info = getattr(self, "_solve_info", None)
        return None if info is None else tuple(float(x) for x in info.cpu().numpy())

    def __getstate__(self):
        return dict(_coeffs=self.get_param_values(), _reg_coeff=self._reg_coeff)

    def __setstate__(self, d):
        self._coeffs, self._reg_coeff = d["_coeffs"], d["_reg_coeff"]
        self._dev_w = None

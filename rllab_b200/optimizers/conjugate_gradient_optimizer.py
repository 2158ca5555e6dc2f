"""ConjugateGradientOptimizer, device-resident (API of rllab/optimizers/conjugate_gradient_optimizer.py:118-296).

update_opt / loss / constraint_val / optimize keep the reference's names and meaning; instead of symbolic Theano
expressions, `loss` names the surrogate kind compiled into the CUDA kernels (b200rl_loss_kl / b200rl_grad /
b200rl_fvp) and `inputs` is the device sample batch.  `optimize` follows :229-296 line by line: loss_before, flat
gradient, krylov.cg with Hessian(KL)-vector products (+reg_coeff*x), initial step size sqrt(2*delta/(x.Hx+1e-8)),
backtracking over backtrack_ratio**k with the accept test `loss < loss_before and kl <= delta`, and the final
reject-and-restore test (note >= there).  All P-vectors stay on the GPU; the host reads two scalars per
line-search trial.  With several GPUs every reduction vector is all-reduced (NCCL) before use.
"""
import numpy as np

from .. import _lib as L
from ..misc import logger


class ConjugateGradientOptimizer(object):
    def __init__(self, cg_iters=10, reg_coeff=1e-5, subsample_factor=1., backtrack_ratio=0.8, max_backtracks=15,
                 accept_violation=False, hvp_approach=None, num_slices=1, residual_tol=1e-10,
                 use_activation_cache=True, precision="f32"):
        if subsample_factor != 1.:
            raise NotImplementedError("subsample_factor < 1 (conjugate_gradient_optimizer.py:235-245) is not on the "
                                      "B200 hot path yet")
        if hvp_approach is not None:
            raise NotImplementedError("only the exact (Perlmutter / Gauss-Newton) Hessian-vector product is built")
        self._cg_iters = cg_iters
        self._reg_coeff = reg_coeff
        self._backtrack_ratio = backtrack_ratio
        self._max_backtracks = max_backtracks
        self._accept_violation = accept_violation
        self._residual_tol = residual_tol
        self._target = None
        self._max_constraint_val = None
        self._constraint_name = None
        self._loss_kind = L.LOSS_TRPO
        self._comm = None
        self._bufs = None
        self._cache = None     # (policy version, batch id) -> (loss, mean_kl, max_kl)
        self._g_key = None     # key for which the `g` buffer holds the flat gradient
        self._hc_key = None    # key for which the batch's activation cache is valid
        self._use_hcache = use_activation_cache and precision == "f32"
        if precision not in ("f32", "f64"):
            raise ValueError("precision must be 'f32' (fast path) or 'f64' (parity mode)")
        self._f64 = precision == "f64"
        self.last_info = {}

    def __getstate__(self):
        return _drop_device_state(self.__dict__, ("_bufs", "_cache", "_g_key", "_hc_key", "_comm"))

    def update_opt(self, loss, target, leq_constraint, inputs=None, extra_inputs=None, constraint_name="constraint",
                   comm=None, *args, **kwargs):
        constraint_term, constraint_value = leq_constraint
        self._loss_kind = loss
        self._target = target
        self._max_constraint_val = constraint_value
        self._constraint_name = constraint_name
        self._comm = comm

    # ---- helpers
    def _buffers(self, P, dev):
        import torch
        if self._bufs is None or self._bufs["g"].numel() != P or self._bufs["g"].device != dev:
            z = lambda n=P: torch.zeros(n, dtype=torch.float64, device=dev)
            self._bufs = dict(g=z(), x=z(), r=z(), p=z(), z=z(), Hx=z(), step=z(), prev=z(), st=z(4), info=z(2),
                              out=z(3))
        return self._bufs

    def _allreduce_out3(self, out):
        if self._comm is not None and self._comm.active:
            self._comm.all_reduce_sum(out[:2])
            self._comm.all_reduce_max(out[2:])

    def _eval(self, batch, want_grad=False):
        """surrogate loss, mean KL, max KL at the target's current parameters (one pass over the batch).
        want_grad: run the gradient pass instead, which yields the same triple for free and leaves the flat
        gradient in the `g` buffer (f_loss + f_grad of cg_opt.py:248,253 fused into one pass)."""
        from .. import ops
        pol = self._target
        key = (pol.version, id(batch), batch.version)
        if self._cache is not None and self._cache[0] == key:
            return self._cache[1]
        b = self._buffers(pol.n_params, batch.device)
        if self._f64:
            ops.update_f64(1 if want_grad else 0, self._loss_kind, pol.theta64, pol.dims, pol.min_std, batch, None,
                           1.0 / batch.B_global, 0.0, 0.0, b["g"] if want_grad else None, b["out"])
            if want_grad:
                if self._comm is not None and self._comm.active:
                    self._comm.all_reduce_sum(b["g"])
                self._g_key = key
        elif want_grad:
            hc = batch.hcache(pol.h1, pol.h2) if self._use_hcache else None
            ops.grad(self._loss_kind, pol.theta32, pol.dims, pol.min_std, batch, 1.0 / batch.B_global, b["g"], b["out"],
                     hc)
            self._hc_key = key if hc is not None else None
            if self._comm is not None and self._comm.active:
                self._comm.all_reduce_sum(b["g"])
            self._g_key = key
        else:
            ops.loss_kl(self._loss_kind, pol.theta32, pol.dims, pol.min_std, batch, 1.0 / batch.B_global, b["out"])
        self._allreduce_out3(b["out"])
        vals = ops.LazyTriple(b["out"])       # pinned-memory readback queued behind the pass; blocks when indexed
        self._cache = (key, vals)
        return vals

    def eval_lazy(self, inputs, want_grad=True):
        """The (loss, mean_kl, max_kl) triple as a lazily read ops.LazyTriple (see FirstOrderOptimizer.eval_lazy)."""
        return self._eval(_lane_batch(inputs), want_grad=want_grad)

    def loss(self, inputs, extra_inputs=None):
        return self._eval(_lane_batch(inputs), want_grad=True)[0]

    def constraint_val(self, inputs, extra_inputs=None):
        return self._eval(_lane_batch(inputs), want_grad=True)[1]

    def optimize(self, inputs, extra_inputs=None, subsample_grouped_inputs=None):
        from .. import ops
        batch = _lane_batch(inputs)
        pol = self._target
        comm = self._comm
        P = pol.n_params
        b = self._buffers(P, batch.device)
        scale = 1.0 / batch.B_global
        world = comm.world_size if (comm is not None and comm.active) else 1
        ar = (lambda t: comm.all_reduce_sum(t)) if world > 1 else (lambda t: t)

        logger.log("computing loss before")
        before = self._eval(batch, want_grad=True)      # read back at the line search, after the CG solve is queued
        logger.log("performing update")
        logger.log("computing descent direction")
        key0 = (pol.version, id(batch), batch.version)
        if self._g_key != key0 and self._f64:
            ops.update_f64(1, self._loss_kind, pol.theta64, pol.dims, pol.min_std, batch, None, scale, 0.0, 0.0,
                           b["g"], None)
            ar(b["g"])
            self._g_key = key0
        if self._g_key != key0:
            hc0 = batch.hcache(pol.h1, pol.h2) if self._use_hcache else None
            ops.grad(self._loss_kind, pol.theta32, pol.dims, pol.min_std, batch, scale, b["g"], None, hc0)
            self._hc_key = key0 if hc0 is not None else None
            ar(b["g"])
        hcache = batch.hcache(pol.h1, pol.h2) if (self._use_hcache and self._hc_key == key0) else None

        def Hx(vec, out):
            if self._f64:
                ops.update_f64(2, self._loss_kind, pol.theta64, pol.dims, pol.min_std, batch, vec, scale,
                               self._reg_coeff, 1.0 / world, out, None)
                ar(out)
                return
            ops.fvp(pol.theta32, pol.dims, pol.min_std, batch, vec, scale, self._reg_coeff, 1.0 / world, out, hcache)
            ar(out)

        ops.cg_init(b["g"], b["x"], b["r"], b["p"], b["st"])
        for _ in range(self._cg_iters):
            Hx(b["p"], b["z"])
            ops.cg_step(b["z"], b["x"], b["r"], b["p"], b["st"], self._residual_tol)
        Hx(b["x"], b["Hx"])
        ops.trpo_step_size(b["x"], b["Hx"], self._max_constraint_val, b["step"], b["info"])
        logger.log("descent direction computed")

        b["prev"].copy_(pol.theta64)
        loss_before = before[0]
        n_iter = 0
        loss = constraint_val = np.nan
        for n_iter, ratio in enumerate(self._backtrack_ratio ** np.arange(self._max_backtracks)):
            ops.axpy_params(b["prev"], b["step"], ratio, pol.theta64, pol.theta32)
            pol.bump_version()
            loss, constraint_val, _ = self._eval(batch)
            if loss < loss_before and constraint_val <= self._max_constraint_val:
                break
        rejected = False
        if (np.isnan(loss) or np.isnan(constraint_val) or loss >= loss_before or
                constraint_val >= self._max_constraint_val) and not self._accept_violation:
            logger.log("Line search condition violated. Rejecting the step!")
            if np.isnan(loss):
                logger.log("Violated because loss is NaN")
            if np.isnan(constraint_val):
                logger.log("Violated because constraint %s is NaN" % self._constraint_name)
            if loss >= loss_before:
                logger.log("Violated because loss not improving")
            if constraint_val >= self._max_constraint_val:
                logger.log("Violated because constraint %s is violated" % self._constraint_name)
            ops.axpy_params(b["prev"], b["step"], 0.0, pol.theta64, pol.theta32)
            pol.bump_version()
            rejected = True
        logger.log("backtrack iters: %d" % n_iter)
        logger.log("computing loss after")
        logger.log("optimization finished")
        self.last_info = dict(loss_before=loss_before, loss=loss, constraint_val=constraint_val, n_iter=n_iter,
                              rejected=rejected)


def _drop_device_state(obj_dict, keys):
    d = dict(obj_dict)
    for k in keys:
        d[k] = None
    return d


def _lane_batch(inputs):
    if hasattr(inputs, "lane_batch"):
        return inputs.lane_batch
    if isinstance(inputs, dict) and "lane_batch" in inputs:
        return inputs["lane_batch"]
    if hasattr(inputs, "obs") and hasattr(inputs, "B_global"):
        return inputs
    raise TypeError("the B200 optimizers take the device sample batch (samples_data or its lane_batch) as `inputs`; "
                    "host arrays would need a host->device copy of the whole batch every call")

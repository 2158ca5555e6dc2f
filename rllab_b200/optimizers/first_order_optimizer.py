"""FirstOrderOptimizer (Adam), device-resident (API of rllab/optimizers/first_order_optimizer.py:15-137).
VPG uses it with batch_size=None, max_epochs=1 (rllab/algos/vpg.py:26-29): one full-batch gradient + one
lasagne.updates.adam step per iteration; Adam moments and the step counter persist across iterations
(they are Theano shared variables created once in update_opt, first_order_optimizer.py:62-65)."""
from .. import _lib as L
from .conjugate_gradient_optimizer import _lane_batch


class FirstOrderOptimizer(object):
    def __init__(self, update_method=None, learning_rate=1e-3, max_epochs=1000, tolerance=1e-6, batch_size=32,
                 callback=None, verbose=False, beta1=0.9, beta2=0.999, epsilon=1e-8, **kwargs):
        if update_method is not None:
            raise NotImplementedError("only lasagne.updates.adam (the reference default) is built")
        self._learning_rate = learning_rate
        self._b1, self._b2, self._eps = beta1, beta2, epsilon
        self._max_epochs = max_epochs
        self._tolerance = tolerance
        self._batch_size = batch_size
        self._callback = callback
        self._verbose = verbose
        self._target = None
        self._loss_kind = L.LOSS_VPG
        self._comm = None
        self._m = self._v = self._g = self._gl = self._gout = self._out = None
        self._t = 0
        self._cache = None
        self._g_key = None

    def update_opt(self, loss, target, inputs=None, extra_inputs=None, gradients=None, comm=None, **kwargs):
        self._loss_kind = loss
        self._target = target
        self._comm = comm
        self._t = 0
        self._m = None

    def _state(self, dev):
        import torch
        P = self._target.n_params
        if self._m is None or self._m.device != dev:
            z = lambda n=P: torch.zeros(n, dtype=torch.float64, device=dev)
            self._gl = z(P + 3)          # [flat gradient | loss, sum KL | max KL]: one collective carries all of it
            self._m, self._v, self._g, self._gout, self._out = z(), z(), self._gl[:P], self._gl[P:], z(3)

    def _eval(self, batch, want_grad=False):
        """(loss, mean_kl, max_kl); want_grad runs the gradient pass, which yields the triple for free and leaves the
        flat gradient in self._g (f_loss + the gradient half of f_opt fused)."""
        from .. import ops
        pol = self._target
        key = (pol.version, id(batch), batch.version)
        if self._cache is not None and self._cache[0] == key:
            return self._cache[1]
        self._state(batch.device)
        active = self._comm is not None and self._comm.active
        fuse = active and self._comm.fuse          # the pass reduces over ranks in its own finalize kernel (peer.cuh)
        if want_grad:
            ops.grad(self._loss_kind, pol.theta32, pol.dims, pol.min_std, batch, self._g, self._gout, fuse=fuse)
            if active:
                self._comm.after_pass(self._gl, pol.n_params + 2)
            self._g_key = key
            src = self._gout
        else:
            ops.loss_kl(self._loss_kind, pol.theta32, pol.dims, pol.min_std, batch, self._out, fuse=fuse)
            if active:
                self._comm.after_pass(self._out, 2)
            src = self._out
        vals = ops.LazyTriple(src)            # pinned-memory readback queued behind the pass; blocks when indexed
        self._cache = (key, vals)
        return vals

    def eval_lazy(self, inputs, want_grad=False):
        """The (loss, mean_kl, max_kl) triple as a lazily read ops.LazyTriple: lets the caller queue the whole
        iteration before the host blocks (the algos record `lambda: triple[0]` with the logger)."""
        return self._eval(_lane_batch(inputs), want_grad=want_grad)

    def loss(self, inputs, extra_inputs=None):
        return self._eval(_lane_batch(inputs), want_grad=True)[0]

    def kl_stats(self, inputs):
        """(mean_kl, max_kl): the f_kl of rllab/algos/vpg.py:100-103, same pass as the loss."""
        v = self._eval(_lane_batch(inputs))
        return v[1], v[2]

    def optimize(self, inputs, extra_inputs=None, callback=None):
        from .. import ops
        batch = _lane_batch(inputs)
        if self._batch_size is not None:
            raise NotImplementedError("mini-batch epochs are not on the B200 hot path (VPG uses batch_size=None)")
        pol = self._target
        self._state(batch.device)
        last = self._eval(batch, want_grad=True)
        for epoch in range(self._max_epochs):
            if self._g_key != (pol.version, id(batch), batch.version):
                active = self._comm is not None and self._comm.active
                ops.grad(self._loss_kind, pol.theta32, pol.dims, pol.min_std, batch, self._g, self._gout,
                         fuse=active and self._comm.fuse)
                if active:
                    self._comm.after_pass(self._gl, pol.n_params + 2)
            self._t += 1
            ops.adam_step(pol.theta64, pol.theta32, self._g, self._m, self._v, self._t, self._learning_rate, self._b1,
                          self._b2, self._eps)
            pol.bump_version()
            new = self._eval(batch)
            if self._callback or callback:
                args = dict(loss=new[0], params=pol.get_param_values(), itr=epoch, elapsed=0.0)
                if self._callback:
                    self._callback(args)
                if callback:
                    callback(**args)
            if epoch + 1 < self._max_epochs:       # the tolerance test only matters if another epoch could follow;
                if abs(last[0] - new[0]) < self._tolerance:   # with max_epochs=1 (VPG) nothing is read back here
                    break
                last = new

    def __getstate__(self):
        d = dict(self.__dict__)
        for k in ("_m", "_v"):
            d[k] = None if d[k] is None else d[k].cpu().numpy()
        for k in ("_g", "_gl", "_gout", "_out"):
            d[k] = None
        d["_cache"] = None
        d["_g_key"] = None
        d["_comm"] = None
        return d

    def __setstate__(self, d):
        import torch
        self.__dict__.update(d)
        if self._m is not None:
            P = len(self._m)
            self._m, self._v = (torch.as_tensor(self.__dict__[k]).cuda() for k in ("_m", "_v"))
            dev = self._m.device
            self._gl = torch.zeros(P + 3, dtype=torch.float64, device=dev)
            self._g, self._gout = self._gl[:P], self._gl[P:]
            self._out = torch.zeros(3, dtype=torch.float64, device=dev)

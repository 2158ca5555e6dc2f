"""Import harness for the *real* reference (rll/rllab) NumPy-side hot path.

TEST INFRASTRUCTURE ONLY.  Used by ``tests/golden/make_golden.py`` (in the build
container, where ``/root/reference`` exists) to pin the oracle restatements in
``oracle/*.py`` against outputs of the reference's own code.  Nothing on the
product path, the GPU tests, ``smoke()`` or ``bench.py`` may import this module:
``/root/reference`` does not exist on the GPU box.

The reference imports a handful of third-party modules at import time that are
absent here: ``path``, ``cached_property``, ``pyprind``,
``joblib.pool.MemmapingPool`` (old spelling), ``theano`` (attribute access only
on the NumPy-side path) and ``_ast.Num``.  We install attribute-only stand-ins so
that the reference's *NumPy* code runs verbatim; anything that would build a
Theano graph still fails loudly.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("RLLAB_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "rllab"))


class _Anything(types.ModuleType):
    """Module whose every attribute is another _Anything (callable, returns itself)."""

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        child = _Anything(self.__name__ + "." + name)
        setattr(self, name, child)
        return child

    def __call__(self, *a, **k):
        return self


def install():
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)

    # --- path.Path (misc/ext.py, config) ---
    if "path" not in sys.modules:
        m = types.ModuleType("path")

        class Path(str):
            pass
        m.Path = Path
        sys.modules["path"] = m

    # --- cached_property (envs/base.py) ---
    if "cached_property" not in sys.modules:
        m = types.ModuleType("cached_property")
        import functools
        m.cached_property = functools.cached_property
        sys.modules["cached_property"] = m

    # --- pyprind (stateful_pool.py, first_order_optimizer.py) ---
    if "pyprind" not in sys.modules:
        m = types.ModuleType("pyprind")

        class ProgBar(object):
            def __init__(self, *a, **k):
                self.active = False

            def update(self, *a, **k):
                pass

            def stop(self):
                pass
        m.ProgBar = ProgBar
        m.prog_bar = lambda it, *a, **k: it
        sys.modules["pyprind"] = m

    # --- joblib.pool.MemmapingPool (old spelling; stateful_pool.py:3) ---
    try:
        import joblib.pool as jp
        if not hasattr(jp, "MemmapingPool") and hasattr(jp, "MemmappingPool"):
            jp.MemmapingPool = jp.MemmappingPool
    except Exception:  # pragma: no cover
        pass

    # --- theano / lasagne: attribute-only ---
    for name in ("theano", "theano.tensor", "theano.tensor.nnet", "theano.tensor.extra_ops",
                 "theano.tensor.signal", "theano.tensor.signal.pool", "theano.sandbox",
                 "theano.sandbox.rng_mrg", "theano.ifelse", "theano.compile", "theano.gradient",
                 "theano.tensor.shared_randomstreams",
                 "lasagne", "lasagne.layers", "lasagne.nonlinearities", "lasagne.init",
                 "lasagne.updates", "lasagne.utils", "lasagne.random"):
        if name not in sys.modules:
            mod = _Anything(name)
            sys.modules[name] = mod
            if "." in name:
                parent, child = name.rsplit(".", 1)
                setattr(sys.modules[parent], child, mod)
    th = sys.modules["theano"]
    th.config.floatX = "float64"  # Theano's default on a plain install

    # --- _ast.Num (removed in py3.12; conjugate_gradient_optimizer.py:10) ---
    import _ast
    import ast
    if not hasattr(_ast, "Num"):
        _ast.Num = ast.Constant
    return True


def import_reference():
    """Returns a namespace with the reference modules that matter for the hot path."""
    install()
    ns = types.SimpleNamespace()
    import rllab.misc.special as special
    import rllab.misc.krylov as krylov
    import rllab.misc.tensor_utils as tensor_utils
    import rllab.sampler.utils as sampler_utils
    import rllab.sampler.base as sampler_base
    import rllab.sampler.parallel_sampler as parallel_sampler
    import rllab.baselines.linear_feature_baseline as lfb
    import rllab.envs.normalized_env as normalized_env
    import rllab.spaces.box as box
    import rllab.algos.util as algo_util
    import rllab.distributions.diagonal_gaussian as diagonal_gaussian
    import rllab.optimizers.conjugate_gradient_optimizer as cg_opt
    import examples.point_env as point_env
    ns.special = special
    ns.krylov = krylov
    ns.tensor_utils = tensor_utils
    ns.sampler_utils = sampler_utils
    ns.sampler_base = sampler_base
    ns.parallel_sampler = parallel_sampler
    ns.lfb = lfb
    ns.normalized_env = normalized_env
    ns.box = box
    ns.algo_util = algo_util
    ns.diagonal_gaussian = diagonal_gaussian
    ns.cg_opt = cg_opt
    ns.point_env = point_env
    return ns

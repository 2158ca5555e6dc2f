"""TRPO (rllab/algos/trpo.py:6-20) = NPO + ConjugateGradientOptimizer."""
from ..optimizers.conjugate_gradient_optimizer import ConjugateGradientOptimizer
from .npo import NPO


class TRPO(NPO):
    def __init__(self, optimizer=None, optimizer_args=None, **kwargs):
        if optimizer is None:
            if optimizer_args is None:
                optimizer_args = dict()
            optimizer = ConjugateGradientOptimizer(**optimizer_args)
        super(TRPO, self).__init__(optimizer=optimizer, **kwargs)


class TNPG(NPO):
    """rllab/algos/tnpg.py: TRPO with max_backtracks=1 (falls out of the same kernels)."""

    def __init__(self, optimizer=None, optimizer_args=None, **kwargs):
        if optimizer is None:
            default_args = dict(max_backtracks=1)
            optimizer_args = default_args if optimizer_args is None else dict(default_args, **optimizer_args)
            optimizer = ConjugateGradientOptimizer(**optimizer_args)
        super(TNPG, self).__init__(optimizer=optimizer, **kwargs)

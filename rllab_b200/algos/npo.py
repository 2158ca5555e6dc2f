"""NPO (rllab/algos/npo.py:10-132): surrogate -mean(lr*adv) under mean KL <= step_size."""
from .. import _lib as L
from ..misc import logger
from .batch_polopt import BatchPolopt


class NPO(BatchPolopt):
    def __init__(self, optimizer=None, optimizer_args=None, step_size=0.01, truncate_local_is_ratio=None, **kwargs):
        if optimizer is None:
            raise NotImplementedError("NPO's default PenaltyLbfgsOptimizer (PPO path) is outside the B200 hot path; "
                                      "use TRPO / pass a ConjugateGradientOptimizer")
        if truncate_local_is_ratio is not None:
            raise NotImplementedError("truncate_local_is_ratio (npo.py:75-76, default off) is not built")
        self.optimizer = optimizer
        self.step_size = step_size
        self.truncate_local_is_ratio = truncate_local_is_ratio
        super(NPO, self).__init__(**kwargs)

    def init_opt(self):
        self.optimizer.update_opt(loss=L.LOSS_TRPO, target=self.policy, leq_constraint=("mean_kl", self.step_size),
                                  inputs=None, constraint_name="mean_kl", comm=getattr(self.sampler, "comm", None))
        return dict()

    def optimize_policy(self, itr, samples_data):
        # npo.py:102-123; lazily read triples (see VPG.optimize_policy): the first host wait is at the line search
        before = self.optimizer.eval_lazy(samples_data)
        self.optimizer.optimize(samples_data)
        after = self.optimizer.eval_lazy(samples_data)
        logger.record_tabular('LossBefore', lambda: before[0])
        logger.record_tabular('LossAfter', lambda: after[0])
        logger.record_tabular('MeanKLBefore', lambda: before[1])
        logger.record_tabular('MeanKL', lambda: after[1])
        logger.record_tabular('dLoss', lambda: before[0] - after[0])
        return dict()

    def get_itr_snapshot(self, itr, samples_data):
        return dict(itr=itr, policy=self.policy, baseline=self.baseline, env=self.env)

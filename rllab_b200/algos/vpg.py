"""VPG (rllab/algos/vpg.py:11-138): surrogate -mean(logp*adv), one full-batch Adam step per iteration."""
from .. import _lib as L
from ..misc import logger
from ..optimizers.first_order_optimizer import FirstOrderOptimizer
from .batch_polopt import BatchPolopt


class VPG(BatchPolopt):
    def __init__(self, env, policy, baseline, optimizer=None, optimizer_args=None, **kwargs):
        if optimizer is None:
            default_args = dict(batch_size=None, max_epochs=1)
            optimizer_args = default_args if optimizer_args is None else dict(default_args, **optimizer_args)
            optimizer = FirstOrderOptimizer(**optimizer_args)
        self.optimizer = optimizer
        self.opt_info = None
        super(VPG, self).__init__(env=env, policy=policy, baseline=baseline, **kwargs)

    def init_opt(self):
        self.optimizer.update_opt(L.LOSS_VPG, target=self.policy, inputs=None,
                                  comm=getattr(self.sampler, "comm", None))
        self.opt_info = dict(f_kl=self.optimizer.kl_stats)

    def optimize_policy(self, itr, samples_data):
        logger.log("optimizing policy")
        # same quantities as vpg.py:110-130; the triples are read back lazily (resolved by logger.dump_tabular) so that
        # the gradient pass, the Adam step and the evaluation pass are all queued before the host blocks
        before = self.optimizer.eval_lazy(samples_data, want_grad=True)
        self.optimizer.optimize(samples_data)
        after = self.optimizer.eval_lazy(samples_data)
        logger.record_tabular("LossBefore", lambda: before[0])
        logger.record_tabular("LossAfter", lambda: after[0])
        logger.record_tabular('MeanKL', lambda: after[1])
        logger.record_tabular('MaxKL', lambda: after[2])

    def get_itr_snapshot(self, itr, samples_data):
        return dict(itr=itr, policy=self.policy, baseline=self.baseline, env=self.env)

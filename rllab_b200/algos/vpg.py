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
        loss_before = self.optimizer.loss(samples_data)
        self.optimizer.optimize(samples_data)
        loss_after = self.optimizer.loss(samples_data)
        logger.record_tabular("LossBefore", loss_before)
        logger.record_tabular("LossAfter", loss_after)
        mean_kl, max_kl = self.opt_info['f_kl'](samples_data)
        logger.record_tabular('MeanKL', mean_kl)
        logger.record_tabular('MaxKL', max_kl)

    def get_itr_snapshot(self, itr, samples_data):
        return dict(itr=itr, policy=self.policy, baseline=self.baseline, env=self.env)

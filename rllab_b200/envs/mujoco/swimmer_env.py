"""SwimmerEnv (rllab/envs/mujoco/swimmer_env.py:10-62); planar restatement in csrc/planar.cuh."""
import numpy as np

from ...misc import logger
from ..lane_env import LaneEnv, require_defaults


class SwimmerEnv(LaneEnv):
    ENV_NAME = "swimmer"

    def __init__(self, ctrl_cost_coeff=1e-2, **kwargs):
        # swimmer_env.py:17-23 + MujocoEnv.__init__(action_noise=0.0, file_path=None, template_args=None)
        require_defaults("SwimmerEnv", dict(kwargs, ctrl_cost_coeff=ctrl_cost_coeff),
                         dict(ctrl_cost_coeff=1e-2, action_noise=0.0, file_path=None, template_args=None))
        self.ctrl_cost_coeff = ctrl_cost_coeff
        super(SwimmerEnv, self).__init__()

    def log_diagnostics(self, paths):
        if len(paths) > 0:
            progs = [path["observations"][-1][-3] - path["observations"][0][-3] for path in paths]
            logger.record_tabular('AverageForwardProgress', np.mean(progs))
            logger.record_tabular('MaxForwardProgress', np.max(progs))
            logger.record_tabular('MinForwardProgress', np.min(progs))
            logger.record_tabular('StdForwardProgress', np.std(progs))
        else:
            for k in ('Average', 'Max', 'Min', 'Std'):
                logger.record_tabular(k + 'ForwardProgress', np.nan)

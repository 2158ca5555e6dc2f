"""HopperEnv (rllab/envs/mujoco/hopper_env.py:19-72); planar restatement in csrc/planar.cuh."""
import numpy as np

from ...misc import logger
from ..lane_env import LaneEnv, require_defaults


class HopperEnv(LaneEnv):
    ENV_NAME = "hopper"

    def __init__(self, alive_coeff=1, ctrl_cost_coeff=0.01, **kwargs):
        # hopper_env.py:27-35 + MujocoEnv.__init__(action_noise=0.0, file_path=None, template_args=None)
        require_defaults("HopperEnv", dict(kwargs, alive_coeff=alive_coeff, ctrl_cost_coeff=ctrl_cost_coeff),
                         dict(alive_coeff=1, ctrl_cost_coeff=0.01, action_noise=0.0, file_path=None,
                              template_args=None))
        self.alive_coeff = alive_coeff
        self.ctrl_cost_coeff = ctrl_cost_coeff
        super(HopperEnv, self).__init__()

    def log_diagnostics(self, paths):
        progs = [path["observations"][-1][-3] - path["observations"][0][-3] for path in paths]
        logger.record_tabular('AverageForwardProgress', np.mean(progs))
        logger.record_tabular('MaxForwardProgress', np.max(progs))
        logger.record_tabular('MinForwardProgress', np.min(progs))
        logger.record_tabular('StdForwardProgress', np.std(progs))

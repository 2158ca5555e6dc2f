"""HopperEnv (rllab/envs/mujoco/hopper_env.py:19-72); planar restatement in csrc/planar.cuh."""
import numpy as np

from ...misc import logger
from ..lane_env import LaneEnv


class HopperEnv(LaneEnv):
    ENV_NAME = "hopper"

    def log_diagnostics(self, paths):
        progs = [path["observations"][-1][-3] - path["observations"][0][-3] for path in paths]
        logger.record_tabular('AverageForwardProgress', np.mean(progs))
        logger.record_tabular('MaxForwardProgress', np.max(progs))
        logger.record_tabular('MinForwardProgress', np.min(progs))
        logger.record_tabular('StdForwardProgress', np.std(progs))

"""SwimmerEnv (rllab/envs/mujoco/swimmer_env.py:10-62); planar restatement in csrc/planar.cuh."""
import numpy as np

from ...misc import logger
from ..lane_env import LaneEnv


class SwimmerEnv(LaneEnv):
    ENV_NAME = "swimmer"

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

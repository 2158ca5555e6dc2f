"""PointEnv (examples/point_env.py:8-30): 2-D point mass, reward -||s||, done near the origin."""
from .lane_env import LaneEnv
from ..spaces import Box
import numpy as np


class PointEnv(LaneEnv):
    ENV_NAME = "point"

    @property
    def observation_space(self):
        return Box(low=-np.inf, high=np.inf, shape=(2,))

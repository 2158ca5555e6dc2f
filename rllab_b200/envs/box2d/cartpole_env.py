"""CartpoleEnv (rllab/envs/box2d/cartpole_env.py:10-56); dynamics restated in csrc/envs.cuh (CartPoleEnvD)."""
from ..lane_env import LaneEnv


class CartpoleEnv(LaneEnv):
    ENV_NAME = "cartpole"

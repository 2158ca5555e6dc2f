"""GymEnv("Pendulum-v0") (rllab/envs/gym_env.py:58-116 over gym==0.7.4): the only gym env on the hot path."""
from .lane_env import LaneEnv


class PendulumEnv(LaneEnv):
    ENV_NAME = "pendulum"
    HORIZON = 200      # gym TimeLimit of Pendulum-v0 [3P]; gym_env.py:104-105 exposes it as env.horizon


def GymEnv(env_name, record_video=False, video_schedule=None, log_dir=None, record_log=False, force_reset=False):
    if env_name != "Pendulum-v0":
        raise NotImplementedError("only GymEnv('Pendulum-v0') is on the B200 hot path (got %r)" % (env_name,))
    return PendulumEnv()

"""GymEnv("Pendulum-v0") (rllab/envs/gym_env.py:58-116 over gym==0.7.4): the only gym env on the hot path."""
from .lane_env import LaneEnv


class PendulumEnv(LaneEnv):
    ENV_NAME = "pendulum"
    HORIZON = 200      # gym TimeLimit of Pendulum-v0 [3P]; gym_env.py:104-105 exposes it as env.horizon


def GymEnv(env_name, record_video=True, video_schedule=None, log_dir=None, record_log=True, force_reset=False):
    """Same signature as gym_env.py:59-60.  The gym Monitor (video / log recording) is outside the hot path: like the
    reference without a snapshot directory (gym_env.py:61-63), monitoring is skipped."""
    if env_name != "Pendulum-v0":
        raise NotImplementedError("only GymEnv('Pendulum-v0') is on the B200 hot path (got %r)" % (env_name,))
    return PendulumEnv()

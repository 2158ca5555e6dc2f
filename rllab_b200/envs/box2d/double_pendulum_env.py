"""DoublePendulumEnv (rllab/envs/box2d/double_pendulum_env.py:11-61); dynamics restated in csrc/envs.cuh
(DoublePendulumEnvD) from models/double_pendulum.xml.mako."""
from ..lane_env import LaneEnv, require_defaults


class DoublePendulumEnv(LaneEnv):
    ENV_NAME = "double_pendulum"

    def __init__(self, **kwargs):
        # double_pendulum_env.py:14-26: frame_skip defaults to 2 (100 ms... per env step: 2 x 0.01 s), link_len 1 unless
        # template_args["noise"]; forwards to Box2DEnv.__init__ (box2d_env.py:30-34)
        require_defaults("DoublePendulumEnv", kwargs, dict(frame_skip=2, position_only=False, obs_noise=0.0,
                                                           action_noise=0.0, template_string=None, template_args=None))
        self.link_len = 1
        super(DoublePendulumEnv, self).__init__()

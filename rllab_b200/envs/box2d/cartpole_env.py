"""CartpoleEnv (rllab/envs/box2d/cartpole_env.py:10-56); dynamics restated in csrc/envs.cuh (CartPoleEnvD)."""
from ..lane_env import LaneEnv, require_defaults


class CartpoleEnv(LaneEnv):
    ENV_NAME = "cartpole"

    def __init__(self, **kwargs):
        # cartpole_env.py:13-24 forwards to Box2DEnv.__init__(model_path, frame_skip=1, position_only=False,
        # obs_noise=0.0, action_noise=0.0, template_string=None, template_args=None) (box2d_env.py:30-34)
        require_defaults("CartpoleEnv", kwargs, dict(frame_skip=1, position_only=False, obs_noise=0.0, action_noise=0.0,
                                                     template_string=None, template_args=None))
        self.max_pole_angle = .2
        self.max_cart_pos = 2.4
        self.max_cart_speed = 4.
        self.max_pole_speed = 4.
        self.reset_range = 0.05
        super(CartpoleEnv, self).__init__()

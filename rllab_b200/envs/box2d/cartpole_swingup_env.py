"""CartpoleSwingupEnv (rllab/envs/box2d/cartpole_swingup_env.py:15-58); same Box2D model as CartpoleEnv, task restated in
csrc/envs.cuh (CartPoleSwingupEnvD)."""
from ..lane_env import LaneEnv, require_defaults


class CartpoleSwingupEnv(LaneEnv):
    ENV_NAME = "cartpole_swingup"

    def __init__(self, **kwargs):
        # cartpole_swingup_env.py:17-22 forwards to Box2DEnv.__init__ (box2d_env.py:30-34)
        require_defaults("CartpoleSwingupEnv", kwargs, dict(frame_skip=1, position_only=False, obs_noise=0.0,
                                                            action_noise=0.0, template_string=None, template_args=None))
        self.max_cart_pos = 3
        self.max_reward_cart_pos = 3
        super(CartpoleSwingupEnv, self).__init__()

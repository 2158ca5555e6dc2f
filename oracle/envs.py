"""Oracle: environment dynamics, lane-batched NumPy restatements (default float64).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Every env is described by a small "lane env" protocol shared with the CUDA path:
    K            number of raw random numbers consumed by one reset
    noise_kind   'uniform' (raw u in [0,1)) or 'normal' (raw N(0,1))
    reset(raw)   raw (K, n) -> state (S, n)
    obs(state)   -> (O, n)
    step(state, u) with u = the action after NormalizedEnv scaling/clipping (A, n)
                 -> (state', reward (n,), done (n,) bool)
    lb, ub       wrapped action bounds used by NormalizedEnv (normalized_env.py:78-92)

PointEnv follows the in-tree examples/point_env.py:16-27 exactly (pinned by golden vectors
generated from the reference itself).  CartPole / Pendulum (and the planar MuJoCo-style
models in oracle/planar.py) restate third-party arithmetic that is absent from
/root/reference: PARITY UNPINNED for those.
"""
import numpy as np


class LaneEnv(object):
    name = None
    kind = -1
    O = A = S = K = 0
    noise_kind = "uniform"
    lb = ub = None

    def __init__(self, dtype=np.float64):
        self.dtype = dtype

    def scale_action(self, a):
        """NormalizedEnv.step, normalized_env.py:81-83: clip(lb + (a+1)*0.5*(ub-lb), lb, ub)."""
        dt = self.dtype
        lb = np.asarray(self.lb, dt).reshape(-1, 1)
        ub = np.asarray(self.ub, dt).reshape(-1, 1)
        a = np.asarray(a, dt)
        scaled = lb + (a + dt(1.0)) * dt(0.5) * (ub - lb)
        return np.clip(scaled, lb, ub)


class PointEnv(LaneEnv):
    """examples/point_env.py:16-27.  reset U(-1,1)^2; s += a; r = -sqrt(x^2+y^2);
    done = |x|<0.01 and |y|<0.01."""
    name, kind = "point", 0
    O, A, S, K = 2, 2, 2, 2
    lb, ub = (-0.1, -0.1), (0.1, 0.1)

    def reset(self, raw):
        dt = self.dtype
        raw = np.asarray(raw, dt)
        return dt(-1.0) + dt(2.0) * raw          # np.random.uniform(-1, 1): low + (high-low)*u

    def obs(self, s):
        return s.copy()

    def step(self, s, u):
        dt = self.dtype
        s2 = (s + u).astype(dt)
        x, y = s2[0], s2[1]
        # python: (x ** 2 + y ** 2) ** 0.5 ; x*x is exactly x**2, and a correctly rounded
        # sqrt equals pow(.,0.5) for the values tested (golden vectors pin this).
        r = -np.sqrt((x * x).astype(dt) + (y * y).astype(dt)).astype(dt)
        done = (np.abs(x) < dt(0.01)) & (np.abs(y) < dt(0.01))
        return s2, r, done


class CartPoleEnv(LaneEnv):
    """Reduced-coordinate restatement of rllab/envs/box2d/cartpole_env.py:13-56 +
    box2d_env.py:119-183 + models/cartpole.xml.mako:3-45 [3P pybox2d: PARITY UNPINNED].

    cart: 4/sqrt12 x 3/sqrt12 box, density 1 -> M = 1.0 ; pole: 0.1 x 1.0 box hinged at its
    bottom edge centre, density 1 -> m = 0.1, COM 0.5 above the hinge, I_com = m (w^2+h^2)/12.
    Box2D gravity (0,-10) [3P default of the parser], dt = 0.05, semi-implicit Euler (Box2D
    integrates v then x).  theta = pole body angle (CCW, 0 = upright).
    state = obs = [x, xdot, theta, thetadot].  force = clip(u, -10, 10) (box2d_env.py:123-124).
    reward (post-step, cartpole_env.py:46-51) = notdone*(10 - (1-cos th) - 1e-5*u^2) with u the
    action handed to env.step (already scaled by NormalizedEnv); done = |x|>2.4 or |th|>0.2.
    """
    name, kind = "cartpole", 1
    O, A, S, K = 4, 1, 4, 4
    lb, ub = (-10.0,), (10.0,)
    M, m, l, g, dt_ = 1.0, 0.1, 0.5, 10.0, 0.05
    I = 0.1 * (0.1 ** 2 + 1.0 ** 2) / 12.0
    bounds = (2.4, 4.0, 0.2, 4.0)
    reset_range = 0.05

    def reset(self, raw):
        dt = self.dtype
        raw = np.asarray(raw, dt)
        b = np.asarray(self.bounds, dt).reshape(4, 1) * dt(self.reset_range)
        return (-b + (dt(2.0) * b) * raw).astype(dt)   # uniform(low, high) = low + (high-low)*u

    def obs(self, s):
        return s.copy()

    def step(self, s, u):
        dt = self.dtype
        x, xd, th, thd = s[0], s[1], s[2], s[3]
        F = np.clip(u[0], dt(-10.0), dt(10.0))
        M, m, l, g, I, h = (dt(v) for v in (self.M, self.m, self.l, self.g, self.I, self.dt_))
        sn, cs = np.sin(th), np.cos(th)
        # (M+m) xdd - m l cos(th) thdd = F - m l sin(th) thd^2
        # -m l cos(th) xdd + (I + m l^2) thdd = m g l sin(th)
        a11 = M + m
        a12 = -m * l * cs
        a22 = I + m * l * l
        b1 = F - m * l * sn * thd * thd
        b2 = m * g * l * sn
        det = a11 * a22 - a12 * a12
        xdd = (a22 * b1 - a12 * b2) / det
        thdd = (a11 * b2 - a12 * b1) / det
        xd2 = xd + h * xdd
        thd2 = thd + h * thdd
        x2 = x + h * xd2
        th2 = th + h * thd2
        s2 = np.stack([x2, xd2, th2, thd2]).astype(dt)
        done = (np.abs(x2) > dt(2.4)) | (np.abs(th2) > dt(0.2))
        notdone = (~done).astype(dt)
        ucost = dt(1e-5) * (u[0] * u[0])
        xcost = dt(1.0) - np.cos(th2)
        r = notdone * dt(10.0) - notdone * xcost - notdone * ucost
        return s2, r.astype(dt), done


class CartPoleSwingupEnv(CartPoleEnv):
    """rllab/envs/box2d/cartpole_swingup_env.py:15-58 on the same Box2D model (cartpole.xml.mako) = the same reduced
    dynamics as CartPoleEnv [3P pybox2d: PARITY UNPINNED]: reset U([-1,-2,pi-1,-3],[1,2,pi+1,3]) (:27-38), done = |x| > 3
    (:54-55), reward (post-step) -100 if done else cos(theta) (:41-51; the -1 branch needs |x| > max_reward_cart_pos = 3 =
    max_cart_pos, i.e. done, so it never fires)."""
    name, kind = "cartpole_swingup", 5

    def reset(self, raw):
        dt = self.dtype
        raw = np.asarray(raw, dt)
        lo = np.asarray([-1.0, -2.0, np.pi - 1.0, -3.0], dt).reshape(4, 1)
        hi = np.asarray([1.0, 2.0, np.pi + 1.0, 3.0], dt).reshape(4, 1)
        return (lo + (hi - lo) * raw).astype(dt)

    def step(self, s, u):
        dt = self.dtype
        s2, _, _ = CartPoleEnv.step(self, s, u)
        done = np.abs(s2[0]) > dt(3.0)
        r = np.where(done, dt(-100.0), np.cos(s2[2])).astype(dt)
        return s2, r, done


class DoublePendulumEnv(LaneEnv):
    """Reduced-coordinate restatement of rllab/envs/box2d/double_pendulum_env.py:11-61 + box2d_env.py:119-183 +
    models/double_pendulum.xml.mako [3P pybox2d: PARITY UNPINNED].

    Two rods (compute_rect_vertices([0,0],[0,-link_len], 0.05): length L = link_len = 1, width 0.1, density 5 -> m = 0.5,
    COM L/2 from the joint, I_com = m (w^2 + L^2)/12); link1 hinged to the static track at the origin, link2 to link1's
    end; absolute body angles th1, th2 (CCW, 0 = hanging down: the rod points along (sin th, -cos th)); Box2D gravity
    (0,-10) [3P default of the parser]; control type="torque" on link_joint_2, ctrllimit +-50: the revolute motor is
    driven at +-1e5 with maxMotorTorque |u| (box2d_env.py:134-144), i.e. torque +u on link2 and -u on link1;
    timestep 0.01, frame_skip 2 (:16), semi-implicit Euler (Box2D integrates v then x).
      M(th) thdd = Q - C - dV:   M11 = I + m lc^2 + m L^2, M22 = I + m lc^2, M12 = m L lc cos(th1 - th2)
      row 1: -u - m L lc sin(th1-th2) w2^2 - (m lc + m L) g sin th1      row 2: +u + m L lc sin(th1-th2) w1^2 - m g lc sin th2
    reset (:31-41): th1, th2 ~ N(0, 0.1), w1, w2 ~ N(0, 0.01).  obs (state tags of the template): sin th1, cos th1, w1,
    sin th2, cos th2, w2.  reward (post-step, :52-58) = -|tip - (0, 2L)| with get_tip_pos (:43-50) = link2.position -
    L (sin th2, cos th2), link2.position = L (sin th1, -cos th1); never done (:60-61)."""
    name, kind = "double_pendulum", 6
    O, A, S, K = 6, 1, 4, 4
    lb, ub = (-50.0,), (50.0,)
    noise_kind = "normal"
    L, m, lc, g, dt_, frame_skip = 1.0, 0.5, 0.5, 10.0, 0.01, 2
    I = 0.5 * (0.1 ** 2 + 1.0 ** 2) / 12.0

    def reset(self, raw):
        dt = self.dtype
        raw = np.asarray(raw, dt)
        stds = np.asarray([0.1, 0.1, 0.01, 0.01], dt).reshape(4, 1)
        return (stds * raw).astype(dt)                     # th1, th2, w1, w2

    def obs(self, s):
        return np.stack([np.sin(s[0]), np.cos(s[0]), s[2], np.sin(s[1]), np.cos(s[1]), s[3]]).astype(self.dtype)

    def step(self, s, u):
        dt = self.dtype
        L, m, lc, g, I, h = (dt(v) for v in (self.L, self.m, self.lc, self.g, self.I, self.dt_))
        th1, th2, w1, w2 = s[0], s[1], s[2], s[3]
        tau = np.clip(u[0], dt(-50.0), dt(50.0))
        m11, m22, mlc = I + m * lc * lc + m * L * L, I + m * lc * lc, m * L * lc
        for _ in range(self.frame_skip):
            sd, cd = np.sin(th1 - th2), np.cos(th1 - th2)
            m12 = mlc * cd
            b1 = -tau - mlc * sd * w2 * w2 - (m * lc + m * L) * g * np.sin(th1)
            b2 = tau + mlc * sd * w1 * w1 - m * g * lc * np.sin(th2)
            det = m11 * m22 - m12 * m12
            a1 = (m22 * b1 - m12 * b2) / det
            a2 = (m11 * b2 - m12 * b1) / det
            w1 = w1 + h * a1
            w2 = w2 + h * a2
            th1 = th1 + h * w1
            th2 = th2 + h * w2
        s2 = np.stack([th1, th2, w1, w2]).astype(dt)
        tx = L * np.sin(th1) - L * np.sin(th2)
        ty = -L * np.cos(th1) - L * np.cos(th2)
        r = -np.sqrt(tx * tx + (ty - dt(2.0) * L) ** 2)
        return s2, r.astype(dt), np.zeros(r.shape, bool)


class PendulumEnv(LaneEnv):
    """gym==0.7.4 Pendulum-v0 (`rllab/envs/gym_env.py:58-116` wraps it; environment.yml:52)
    [3P gym: PARITY UNPINNED].  max_speed 8, max_torque 2, dt .05, g 10, m 1, l 1.
    state=[th, thdot]; obs=[cos th, sin th, thdot]; reset th~U(-pi,pi), thdot~U(-1,1);
    cost = angle_normalize(th)^2 + .1 thdot^2 + .001 u^2 (pre-step state, clipped u); never done."""
    name, kind = "pendulum", 2
    O, A, S, K = 3, 1, 2, 2
    lb, ub = (-2.0,), (2.0,)

    def reset(self, raw):
        dt = self.dtype
        raw = np.asarray(raw, dt)
        high = np.asarray([np.pi, 1.0], dt).reshape(2, 1)
        return (-high + (dt(2.0) * high) * raw).astype(dt)

    def obs(self, s):
        return np.stack([np.cos(s[0]), np.sin(s[0]), s[1]]).astype(self.dtype)

    def step(self, s, u):
        dt = self.dtype
        th, thd = s[0], s[1]
        uu = np.clip(u[0], dt(-2.0), dt(2.0))
        two_pi = dt(2.0 * np.pi)
        pi = dt(np.pi)
        an = np.mod(th + pi, two_pi) - pi            # angle_normalize
        cost = an * an + dt(0.1) * thd * thd + dt(0.001) * (uu * uu)
        # -3g/(2l) sin(th+pi) + 3/(m l^2) u  with g=10,l=1,m=1
        newthd = thd + (dt(-15.0) * np.sin(th + pi) + dt(3.0) * uu) * dt(0.05)
        newth = th + newthd * dt(0.05)
        newthd = np.clip(newthd, dt(-8.0), dt(8.0))
        s2 = np.stack([newth, newthd]).astype(dt)
        done = np.zeros(th.shape, dtype=bool)
        return s2, (-cost).astype(dt), done


def make(name, dtype=np.float64):
    name = name.lower()
    if name in ("point", "pointenv"):
        return PointEnv(dtype)
    if name in ("cartpole", "cartpoleenv"):
        return CartPoleEnv(dtype)
    if name in ("cartpole_swingup", "cartpoleswingupenv"):
        return CartPoleSwingupEnv(dtype)
    if name in ("double_pendulum", "doublependulumenv"):
        return DoublePendulumEnv(dtype)
    if name in ("pendulum", "pendulum-v0"):
        return PendulumEnv(dtype)
    if name in ("swimmer", "hopper"):
        from . import planar
        return planar.make(name, dtype)
    raise ValueError(name)

"""Oracle: planar articulated-body restatements of rllab's MuJoCo Swimmer and Hopper (lane-batched NumPy, float64
by default).  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Reference call sites: rllab/envs/mujoco/mujoco_env.py:109-132,184-191 (reset: qpos+0.01 N, qvel+0.1 N; step:
frame_skip x mj_step, mj_forward), swimmer_env.py:25-45, hopper_env.py:38-61, mujoco_py/mjcore.py:58-81 (subtree
COM velocity from BODY-ORIGIN velocities), models vendor/mujoco_models/swimmer.xml, hopper.xml.

The arithmetic itself lives in the closed MuJoCo 1.31 binary, which is absent: **PARITY UNPINNED** (SURVEY 8c).
What is restated here is MuJoCo's published model (generalised coordinates, M(q) qacc + c = tau, semi-implicit Euler
or RK4, inertia-box fluid forces, soft constraints with solref/solimp reference acceleration solved by projected
Gauss-Seidel) specialised to the two planar serial chains:

  q = [tX, tY (order per model), hinge_0 .. hinge_{n-1}],  body i absolute angle phi_i = sum_{k<=i} s_k q_hinge_k
  hinge position h_0 = (tX, tY), h_i = h_{i-1} + R(phi_{i-1}) a_i ; COM p_i = h_i + R(phi_i) c_i
  M = sum_i m_i J_i^T J_i + I_i w_i w_i^T + armature ;  bias = sum_i m_i J_i^T (centripetal acc. of p_i)
  passive: joint damping, fluid (swimmer);  gravity (hopper);  actuation: torque = clip(ctrl) on the driven hinges
  constraints: joint limits, foot-sphere/ground contacts (normal + 1 tangential friction row), solved with a
  fixed number of PGS sweeps on  (A + R) f = aref - J a0,  A = J M^-1 J^T, R = (1-d)/d diag(A).
The CUDA kernels (rllab_b200/csrc/planar.cuh) implement exactly this; tests compare them with tolerance.
"""
import numpy as np

from .envs import LaneEnv

PGS_SWEEPS = 8


def capsule(r, L, rho=1000.0):
    """mass, inertia about COM perpendicular to the axis, inertia about the axis (MuJoCo capsule formulas)."""
    mc = rho * np.pi * r * r * L
    mh = rho * (2.0 / 3.0) * np.pi * r ** 3
    m = mc + 2 * mh
    Ip = mc * (r * r / 4 + L * L / 12) + 2 * mh * (2 * r * r / 5 + L * L / 4 + 3 * L * r / 8)
    Ia = mc * r * r / 2 + 2 * mh * (2 * r * r / 5)
    return m, Ip, Ia


class Model(object):
    pass


def swimmer_model():
    m = Model()
    m.name = "swimmer"
    m.n = 3                                   # bodies / hinges
    m.iX, m.iY = 0, 1                         # q index of the X / Y translation
    m.sgn = [1.0, 1.0, 1.0]                   # hinge sign (axis +z -> CCW)
    m.a = [(0.0, 0.0), (0.5, 0.0), (-1.0, 0.0)]      # hinge anchor in the parent's hinge frame
    m.c = [(1.0, 0.0), (-0.5, 0.0), (-0.5, 0.0)]     # COM in the body's hinge frame
    m.bo = [(0.0, 0.0), (0.0, 0.0), (0.0, 0.0)]      # MuJoCo body-frame origin in the hinge frame (comvel quirk)
    caps = [capsule(0.1, 1.0)] * 3
    m.mass = [c[0] for c in caps]
    m.Ip = [c[1] for c in caps]
    m.Ia = [c[2] for c in caps]
    m.long_axis = [(1.0, 0.0)] * 3            # capsule axis in the body frame (for the fluid inertia box)
    m.armature = [0.0] * 5
    m.damping = [0.0] * 5
    m.gravity = (0.0, 0.0)
    m.density, m.viscosity = 4000.0, 0.1
    m.act = [1, 2]                            # actuated hinges (rot2, rot3)
    m.ctrl_lim = 50.0
    m.limits = [None, (-100.0 * np.pi / 180, 100.0 * np.pi / 180), (-100.0 * np.pi / 180, 100.0 * np.pi / 180)]
    m.lim_solref, m.lim_solimp = (0.02, 1.0), (0.9, 0.95, 0.001)
    m.contacts = []
    m.dt, m.frame_skip, m.rk4 = 0.001, 50, False
    m.q0 = [0.0] * 5
    return m


def hopper_model():
    m = Model()
    m.name = "hopper"
    m.n = 4
    m.iX, m.iY = 1, 0                         # q = [rootz, rootx, rooty, thigh, leg, foot]; plane (X=x, Y=z)
    m.sgn = [-1.0, 1.0, 1.0, 1.0]             # rooty about +y = clockwise in (x,z); leg joints about -y = CCW
    m.a = [(0.0, 0.0), (0.0, -0.2), (0.0, -0.45), (0.0, -0.5)]
    m.c = [(0.0, 0.0), (0.0, -0.225), (0.0, -0.25), (0.065, 0.0)]
    m.bo = [(0.0, 0.0), (0.0, 0.0), (0.0, -0.25), (0.065, 0.0)]
    caps = [capsule(0.05, 0.4), capsule(0.05, 0.45), capsule(0.04, 0.5), capsule(0.06, 0.39)]
    m.mass = [c[0] for c in caps]
    m.Ip = [c[1] for c in caps]
    m.Ia = [c[2] for c in caps]
    m.long_axis = [(0.0, 1.0), (0.0, 1.0), (0.0, 1.0), (1.0, 0.0)]
    m.armature = [0.0, 0.0, 0.0, 1.0, 1.0, 1.0]
    m.damping = [0.0, 0.0, 0.0, 1.0, 1.0, 1.0]
    m.gravity = (0.0, -9.81)
    m.density, m.viscosity = 0.0, 0.0
    m.act = [1, 2, 3]
    m.ctrl_lim = 200.0
    d2r = np.pi / 180
    m.limits = [None, (-150 * d2r, 0.0), (-150 * d2r, 0.0), (-45 * d2r, 45 * d2r)]
    m.lim_solref, m.lim_solimp = (0.02, 1.0), (0.9, 0.95, 0.001)
    # foot capsule end spheres (hinge frame of body 3), radius, friction, margin, solref, solimp
    m.contacts = [dict(body=3, e=(-0.13, 0.0), r=0.06), dict(body=3, e=(0.26, 0.0), r=0.06)]
    m.mu, m.margin = 2.0, 0.001
    m.con_solref, m.con_solimp = (0.02, 1.0), (0.8, 0.8, 0.01)
    m.dt, m.frame_skip, m.rk4 = 0.02, 1, True
    m.q0 = [1.25, 0.0, 0.0, 0.0, 0.0, 0.0]
    return m


def _rot(cs, sn, v):
    return (cs * v[0] - sn * v[1], sn * v[0] + cs * v[1])


def _impedance(solimp, r):
    d0, d1, w = solimp
    x = np.minimum(np.abs(r) / w, 1.0)
    return d0 + (d1 - d0) * x


def _kb(solref, solimp):
    tc, dr = solref
    dmax = max(solimp[0], solimp[1])
    b = 2.0 / (dmax * tc)
    kk = 1.0 / (dmax * dmax * tc * tc * dr * dr)
    return kk, b


def dynamics(m, q, v, ctrl, dt=np.float64):
    """qacc (nv,N), qfrc_constraint (nv,N), kin dict.  q, v: lists/arrays (nv, N); ctrl (nu, N)."""
    n, nv = m.n, m.n + 2
    N = q[0].shape[0]
    z = np.zeros(N, dt)
    one = np.ones(N, dt)
    # ---- kinematics
    phi, om = [], []
    acc_p, acc_w = z, z
    for i in range(n):
        acc_p = acc_p + dt(m.sgn[i]) * q[2 + i]
        acc_w = acc_w + dt(m.sgn[i]) * v[2 + i]
        phi.append(acc_p)
        om.append(acc_w)
    cs = [np.cos(p) for p in phi]
    sn = [np.sin(p) for p in phi]
    h = [(q[m.iX], q[m.iY])]
    hd = [(v[m.iX], v[m.iY])]
    hdd = [(z, z)]
    for i in range(1, n):
        ra = _rot(cs[i - 1], sn[i - 1], (dt(m.a[i][0]), dt(m.a[i][1])))
        h.append((h[i - 1][0] + ra[0], h[i - 1][1] + ra[1]))
        hd.append((hd[i - 1][0] - om[i - 1] * ra[1], hd[i - 1][1] + om[i - 1] * ra[0]))
        w2 = om[i - 1] * om[i - 1]
        hdd.append((hdd[i - 1][0] - w2 * ra[0], hdd[i - 1][1] - w2 * ra[1]))
    p, pd, pdd = [], [], []
    for i in range(n):
        rc = _rot(cs[i], sn[i], (dt(m.c[i][0]), dt(m.c[i][1])))
        p.append((h[i][0] + rc[0], h[i][1] + rc[1]))
        pd.append((hd[i][0] - om[i] * rc[1], hd[i][1] + om[i] * rc[0]))
        w2 = om[i] * om[i]
        pdd.append((hdd[i][0] - w2 * rc[0], hdd[i][1] - w2 * rc[1]))

    def point_jac(body, pt):
        """Jacobian rows (JX, JY), each a list of nv arrays, of a point fixed to `body`."""
        JX = [z] * nv
        JY = [z] * nv
        JX[m.iX] = one
        JY[m.iY] = one
        for k in range(body + 1):
            JX[2 + k] = -dt(m.sgn[k]) * (pt[1] - h[k][1])
            JY[2 + k] = dt(m.sgn[k]) * (pt[0] - h[k][0])
        return JX, JY

    # ---- mass matrix, bias, applied forces
    M = [[z for _ in range(nv)] for _ in range(nv)]
    tau = [z for _ in range(nv)]
    for i in range(n):
        JX, JY = point_jac(i, p[i])
        mi, Ii = dt(m.mass[i]), dt(m.Ip[i])
        fX = mi * dt(m.gravity[0]) - mi * pdd[i][0]
        fY = mi * dt(m.gravity[1]) - mi * pdd[i][1]
        tq = z
        if m.density > 0 or m.viscosity > 0:
            la = _rot(cs[i], sn[i], (dt(m.long_axis[i][0]), dt(m.long_axis[i][1])))   # world long axis
            vl = pd[i][0] * la[0] + pd[i][1] * la[1]
            vp = -pd[i][0] * la[1] + pd[i][1] * la[0]
            bl = dt(np.sqrt(6.0 * (2 * m.Ip[i] - m.Ia[i]) / m.mass[i]))
            bp = dt(np.sqrt(6.0 * m.Ia[i] / m.mass[i]))
            diam = (bl + 2 * bp) / dt(3.0)
            rho, beta = dt(m.density), dt(m.viscosity)
            Fl = -dt(0.5) * rho * bp * bp * np.abs(vl) * vl - dt(3 * np.pi) * beta * diam * vl
            Fp = -dt(0.5) * rho * bl * bp * np.abs(vp) * vp - dt(3 * np.pi) * beta * diam * vp
            fX = fX + Fl * la[0] - Fp * la[1]
            fY = fY + Fl * la[1] + Fp * la[0]
            tq = -rho * bp * (bl ** 4 + bp ** 4) / dt(64.0) * np.abs(om[i]) * om[i] - dt(np.pi) * beta * diam ** 3 * om[i]
        wv = [z] * nv
        for k in range(i + 1):
            wv[2 + k] = dt(m.sgn[k]) * one
        for r in range(nv):
            tau[r] = tau[r] + JX[r] * fX + JY[r] * fY + wv[r] * tq
            for c_ in range(r, nv):
                M[r][c_] = M[r][c_] + mi * (JX[r] * JX[c_] + JY[r] * JY[c_]) + Ii * wv[r] * wv[c_]
    for r in range(nv):
        M[r][r] = M[r][r] + dt(m.armature[r])
        tau[r] = tau[r] - dt(m.damping[r]) * v[r]
        for c_ in range(r):
            M[r][c_] = M[c_][r]
    for j, hk in enumerate(m.act):
        tau[2 + hk] = tau[2 + hk] + np.clip(ctrl[j], -dt(m.ctrl_lim), dt(m.ctrl_lim))

    # ---- Cholesky M = L L^T (lane-wise)
    Lc = [[z for _ in range(nv)] for _ in range(nv)]
    for r in range(nv):
        for c_ in range(r + 1):
            s = M[r][c_]
            for k in range(c_):
                s = s - Lc[r][k] * Lc[c_][k]
            Lc[r][c_] = np.sqrt(s) if r == c_ else s / Lc[c_][c_]

    def solve(b):
        y = [None] * nv
        for r in range(nv):
            s = b[r]
            for k in range(r):
                s = s - Lc[r][k] * y[k]
            y[r] = s / Lc[r][r]
        x = [None] * nv
        for r in range(nv - 1, -1, -1):
            s = y[r]
            for k in range(r + 1, nv):
                s = s - Lc[k][r] * x[k]
            x[r] = s / Lc[r][r]
        return x

    a0 = solve(tau)

    # ---- constraints: rows (J, aref, d, kind, partner, mu)
    rows = []
    kl, bl_ = _kb(m.lim_solref, m.lim_solimp)
    for k in range(n):
        if m.limits[k] is None:
            continue
        lo, hi = m.limits[k]
        for side, sgn_ in ((lo, 1.0), (hi, -1.0)):
            r_ = dt(sgn_) * (q[2 + k] - dt(side))            # >= 0 when inside
            J = [z] * nv
            J[2 + k] = dt(sgn_) * one
            d = _impedance(m.lim_solimp, r_)
            aref = -dt(bl_) * (dt(sgn_) * v[2 + k]) - dt(kl) * d * r_
            rows.append(dict(J=J, aref=aref, d=d, active=(r_ < 0), normal=None))
    if m.contacts:
        kc, bc = _kb(m.con_solref, m.con_solimp)
        for cdef in m.contacts:
            bi = cdef["body"]
            e = _rot(cs[bi], sn[bi], (dt(cdef["e"][0]), dt(cdef["e"][1])))
            sc = (h[bi][0] + e[0], h[bi][1] + e[1])
            dist = sc[1] - dt(cdef["r"])
            pt = (sc[0], sc[1] - dt(cdef["r"]))
            JX, JY = point_jac(bi, pt)
            r_ = dist - dt(m.margin)
            d = _impedance(m.con_solimp, r_)
            vn = sum(JY[k] * v[k] for k in range(nv))
            vt = sum(JX[k] * v[k] for k in range(nv))
            active = r_ < 0
            rows.append(dict(J=JY, aref=-dt(bc) * vn - dt(kc) * d * r_, d=d, active=active, normal=None))
            rows.append(dict(J=JX, aref=-dt(bc) * vt, d=d, active=active, normal=len(rows) - 1))
    nc = len(rows)
    qfc = [z for _ in range(nv)]
    if nc > 0:
        MiJ = [solve(rw["J"]) for rw in rows]                    # M^-1 J_i^T
        A = [[sum(rows[i]["J"][k] * MiJ[j][k] for k in range(nv)) for j in range(nc)] for i in range(nc)]
        rhs = [rows[i]["aref"] - sum(rows[i]["J"][k] * a0[k] for k in range(nv)) for i in range(nc)]
        Rr = [(dt(1.0) - rows[i]["d"]) / rows[i]["d"] * A[i][i] for i in range(nc)]
        f = [z for _ in range(nc)]
        for _ in range(PGS_SWEEPS):
            for i in range(nc):
                s = rhs[i] - Rr[i] * f[i]
                for j in range(nc):
                    s = s - A[i][j] * f[j]
                fi = f[i] + s / (A[i][i] + Rr[i])
                if rows[i]["normal"] is None:
                    fi = np.maximum(fi, 0.0)
                else:
                    lim = dt(m.mu) * f[rows[i]["normal"]]
                    fi = np.clip(fi, -lim, lim)
                f[i] = np.where(rows[i]["active"], fi, z)
        for i in range(nc):
            for k in range(nv):
                qfc[k] = qfc[k] + rows[i]["J"][k] * f[i]
    acc = solve([tau[k] + qfc[k] for k in range(nv)]) if nc > 0 else a0
    # ---- subtree COM and the reference's body-origin "COM velocity" (mjcore.py:58-81)
    mt = sum(m.mass)
    comX = sum(dt(m.mass[i]) * p[i][0] for i in range(n)) / dt(mt)
    comY = sum(dt(m.mass[i]) * p[i][1] for i in range(n)) / dt(mt)
    cvX = z
    for i in range(n):
        ro = _rot(cs[i], sn[i], (dt(m.bo[i][0]), dt(m.bo[i][1])))
        cvX = cvX + dt(m.mass[i]) * (hd[i][0] - om[i] * ro[1])
    cvX = cvX / dt(mt)
    return acc, qfc, dict(comX=comX, comY=comY, comvelX=cvX)


def integrate(m, q, v, ctrl, dt=np.float64):
    """One env step: frame_skip x (semi-implicit Euler | RK4).  Returns new (q, v) as (nv,N) arrays."""
    nv = m.n + 2
    h = dt(m.dt)
    q = [np.asarray(x, dt) for x in q]
    v = [np.asarray(x, dt) for x in v]
    for _ in range(m.frame_skip):
        if not m.rk4:
            a, _, _ = dynamics(m, q, v, ctrl, dt)
            v = [v[k] + h * a[k] for k in range(nv)]
            q = [q[k] + h * v[k] for k in range(nv)]
        else:
            k1v, _, _ = dynamics(m, q, v, ctrl, dt)
            k1q = v
            q2 = [q[k] + dt(0.5) * h * k1q[k] for k in range(nv)]
            v2 = [v[k] + dt(0.5) * h * k1v[k] for k in range(nv)]
            k2v, _, _ = dynamics(m, q2, v2, ctrl, dt)
            q3 = [q[k] + dt(0.5) * h * v2[k] for k in range(nv)]
            v3 = [v[k] + dt(0.5) * h * k2v[k] for k in range(nv)]
            k3v, _, _ = dynamics(m, q3, v3, ctrl, dt)
            q4 = [q[k] + h * v3[k] for k in range(nv)]
            v4 = [v[k] + h * k3v[k] for k in range(nv)]
            k4v, _, _ = dynamics(m, q4, v4, ctrl, dt)
            s6 = h / dt(6.0)
            q = [q[k] + s6 * (k1q[k] + dt(2.0) * v2[k] + dt(2.0) * v3[k] + v4[k]) for k in range(nv)]
            v = [v[k] + s6 * (k1v[k] + dt(2.0) * k2v[k] + dt(2.0) * k3v[k] + k4v[k]) for k in range(nv)]
    return np.stack(q), np.stack(v)


class SwimmerEnv(LaneEnv):
    """rllab/envs/mujoco/swimmer_env.py:10-45.  state = [qpos(5), qvel(5)]; obs = [qpos, qvel, com(3)];
    reward = comvel_x - 0.5*1e-2*sum((a/50)^2); never done."""
    name, kind = "swimmer", 3
    O, A, S, K = 13, 2, 10, 10
    noise_kind = "normal"
    lb, ub = (-50.0, -50.0), (50.0, 50.0)

    def __init__(self, dtype=np.float64):
        LaneEnv.__init__(self, dtype)
        self.m = swimmer_model()

    def reset(self, raw):
        dt = self.dtype
        raw = np.asarray(raw, dt)
        q = np.asarray(self.m.q0, dt).reshape(-1, 1) + dt(0.01) * raw[:5]
        v = dt(0.1) * raw[5:10]
        return np.concatenate([q, v]).astype(dt)

    def obs(self, s):
        dt = self.dtype
        _, _, kin = dynamics(self.m, list(s[:5]), list(s[5:10]), np.zeros((2, s.shape[1]), dt), dt)
        return np.concatenate([s[:10], np.stack([kin["comX"], kin["comY"], np.zeros_like(kin["comX"])])]).astype(dt)

    def step(self, s, u):
        dt = self.dtype
        q, v = integrate(self.m, list(s[:5]), list(s[5:10]), u, dt)
        _, _, kin = dynamics(self.m, list(q), list(v), u, dt)
        ctrl_cost = dt(0.5 * 1e-2) * ((u[0] / dt(50.0)) ** 2 + (u[1] / dt(50.0)) ** 2)
        r = kin["comvelX"] - ctrl_cost
        s2 = np.concatenate([q, v]).astype(dt)
        return s2, r.astype(dt), np.zeros(s.shape[1], bool)


class HopperEnv(LaneEnv):
    """rllab/envs/mujoco/hopper_env.py:19-61.  state = [qpos(6), qvel(6), ctrl(3), qfrc_constraint(6), comX, comY]
    (the last 8 cache mj_forward's outputs at the current state, as the CUDA env does; obs recomputes them here);
    obs = [q0, q2..q5, clip(qvel,+-10), clip(qfrc_constraint,+-10), com(3)];
    reward = comvel_x + 1 - 0.5*0.01*sum((a/200)^2); done = not(finite and |state[3:]|<100 and z>.7 and |pitch|<.2)."""
    name, kind = "hopper", 4
    O, A, S, K = 20, 3, 23, 12
    noise_kind = "normal"
    lb, ub = (-200.0,) * 3, (200.0,) * 3

    def __init__(self, dtype=np.float64):
        LaneEnv.__init__(self, dtype)
        self.m = hopper_model()

    def reset(self, raw):
        dt = self.dtype
        raw = np.asarray(raw, dt)
        q = np.asarray(self.m.q0, dt).reshape(-1, 1) + dt(0.01) * raw[:6]
        v = dt(0.1) * raw[6:12]
        s = np.concatenate([q, v, np.zeros((3, raw.shape[1]), dt)]).astype(dt)
        return self._with_cache(s)

    def _with_cache(self, s15):
        dt = self.dtype
        _, qfc, kin = dynamics(self.m, list(s15[:6]), list(s15[6:12]), s15[12:15], dt)
        return np.concatenate([s15[:15], np.stack(qfc), np.stack([kin["comX"], kin["comY"]])]).astype(dt)

    def obs(self, s):
        dt = self.dtype
        _, qfc, kin = dynamics(self.m, list(s[:6]), list(s[6:12]), s[12:15], dt)
        qfc = np.stack(qfc)
        return np.concatenate([s[0:1], s[2:6], np.clip(s[6:12], -10, 10), np.clip(qfc, -10, 10),
                               np.stack([kin["comX"], np.zeros_like(kin["comX"]), kin["comY"]])]).astype(dt)

    def step(self, s, u):
        dt = self.dtype
        q, v = integrate(self.m, list(s[:6]), list(s[6:12]), u, dt)
        _, _, kin = dynamics(self.m, list(q), list(v), u, dt)
        cost = dt(0.5 * 0.01) * sum((u[k] / dt(200.0)) ** 2 for k in range(3))
        r = kin["comvelX"] + dt(1.0) - cost
        st = np.concatenate([q, v])
        notdone = np.isfinite(st).all(axis=0) & (np.abs(st[3:]) < 100).all(axis=0) & (st[0] > 0.7) & \
            (np.abs(st[2]) < 0.2)
        s2 = self._with_cache(np.concatenate([q, v, np.asarray(u, dt)]).astype(dt))
        return s2, r.astype(dt), ~notdone


def make(name, dtype=np.float64):
    return SwimmerEnv(dtype) if name == "swimmer" else HopperEnv(dtype)

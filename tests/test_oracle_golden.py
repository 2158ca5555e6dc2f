"""CPU: the oracle restatements against golden vectors produced by the reference's own code
(tests/golden/make_golden.py ran /root/reference verbatim in the build container)."""
import numpy as np

from oracle import envs as E
from oracle import optim as OPT
from oracle import policy as P
from oracle import sampler as S


def _traj(g):
    return {k[len("ps_in_"):]: v for k, v in g.items() if k.startswith("ps_in_") and k != "ps_in_coeffs_prev"}


def test_process_samples_matches_reference(golden):
    g = golden
    traj = _traj(g)
    for tag, coeffs in (("a", None), ("b", g["ps_in_coeffs_prev"]), ("c", g["ps_in_coeffs_prev"])):
        disc, lam, center, positive = g["ps_%s_cfg" % tag]
        out = S.process_samples_lanes(traj, coeffs, disc, lam, bool(center), bool(positive))
        np.testing.assert_allclose(out["adv"], g["ps_%s_adv" % tag], rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(out["ret"], g["ps_%s_ret" % tag], rtol=1e-12, atol=1e-12)
        for key in ("AverageDiscountedReturn", "AverageReturn", "ExplainedVariance", "Entropy", "Perplexity",
                    "StdReturn", "MaxReturn", "MinReturn"):
            np.testing.assert_allclose(out["stats"][key], g["ps_%s_%s" % (tag, key)], rtol=1e-10, err_msg=key)
        assert out["stats"]["NumTrajs"] == int(g["ps_%s_NumTrajs" % tag])       # integer: exact
        fit = S.lfb_fit_lanes(traj["obs"], traj["tstep"], out["ret"])
        np.testing.assert_allclose(fit, g["ps_%s_fit" % tag], rtol=1e-7, atol=1e-9)


def test_cg_matches_reference(golden):
    g = golden
    A, b = g["cg_A"], g["cg_b"]
    np.testing.assert_allclose(OPT.cg(lambda x: A @ x, b.copy(), 10), g["cg_x10"], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(OPT.cg(lambda x: A @ x, b.copy(), 3), g["cg_x3"], rtol=1e-12, atol=1e-14)


def test_trpo_optimize_matches_reference(golden):
    g = golden
    dims = P.Dims(3, (8, 8), 2)
    batch = dict(obs=g["opt_in_obs"], adv=g["opt_in_adv"], old_mean=g["opt_in_old_mean"],
                 old_log_std=g["opt_in_old_log_std"], actions=g["opt_in_actions"])
    theta0 = g["opt_theta0"]
    seen = {}
    for tag in ("acc", "small", "rej"):
        step_size, scale = g["opt_%s_cfg" % tag]
        f_loss = lambda th: P.surr_loss_trpo(th, batch, dims)
        f_grad = lambda th: scale * P.grad_surr(th, batch, dims, "trpo")
        f_lc = lambda th: (P.surr_loss_trpo(th, batch, dims), P.kl_stats(th, batch, dims)[0])
        f_Hx = lambda th, x: P.fvp(theta0, batch, x, dims, 1e-5)
        th, info = OPT.trpo_optimize(f_loss, f_grad, f_lc, f_Hx, theta0, step_size)
        np.testing.assert_allclose(th, g["opt_%s_theta" % tag], rtol=1e-11, atol=1e-13)
        seen[tag] = info
    assert seen["rej"]["rejected"] and np.array_equal(g["opt_rej_theta"], theta0)
    assert not seen["acc"]["rejected"] and not np.array_equal(g["opt_acc_theta"], theta0)


def test_diagonal_gaussian_matches_reference(golden):
    g = golden
    np.testing.assert_allclose(P.kl(g["dg_om"], g["dg_ol"], g["dg_nm"], g["dg_nl"]), g["dg_kl"], rtol=1e-13)
    np.testing.assert_allclose(P.log_likelihood(g["dg_xs"], g["dg_nm"], g["dg_nl"]), g["dg_ll"], rtol=1e-13)
    np.testing.assert_allclose(P.entropy(g["dg_nl"]), g["dg_ent"], rtol=1e-13)
    # docs/user/experiments.rst:88 -- Entropy 1.41894 at log_std = 0, A = 1
    assert abs(float(P.entropy(np.zeros(1))) - 1.41894) < 1e-5


class _ReplayDims(object):
    pass


def _replay_point(s0, actions, T):
    """Drive the oracle lane rollout with pre-drawn actions (policy with zero weights and log_std=-inf
    is awkward; instead step the env restatement directly, mirroring sampler/utils.py:18-29)."""
    env = E.make("point")
    s = np.asarray(s0, np.float64).reshape(2, 1)
    obs, rew = [], []
    for t in range(T):
        obs.append(env.obs(s)[:, 0])
        s, r, d = env.step(s, env.scale_action(actions[t].reshape(2, 1)))
        rew.append(r[0])
        if d[0]:
            break
    return np.array(obs), np.array(rew)


def test_point_env_matches_reference_bit_exact(golden):
    g = golden
    obs, rew = _replay_point(g["pt_f64_s0"], g["pt_f64_actions"], 40)
    assert np.array_equal(obs, g["pt_f64_obs"])            # bit-exact in float64
    assert np.array_equal(rew, g["pt_f64_rew"])
    obs, rew = _replay_point([0.05, -0.03], g["pt_done_actions"], 10)
    assert len(rew) == int(g["pt_done_len"])               # early termination index: exact
    assert np.array_equal(obs, g["pt_done_obs"]) and np.array_equal(rew, g["pt_done_rew"])


def test_truncate_paths_matches_reference(golden):
    g = golden
    for ms in (130, 150, 1, 249, 250, 400):
        assert S.truncate_paths_lengths(g["tr_lens"], ms) == list(g["tr_%d" % ms])
    # tests/test_sampler.py:4-32 known answer
    assert S.truncate_paths_lengths([100, 50], 130) == [100, 30]


def test_misc_matches_reference(golden):
    g = golden
    np.testing.assert_allclose(S.discount_cumsum(g["misc_x"], 0.97), g["misc_dcs"], rtol=1e-12)
    np.testing.assert_allclose(S.explained_variance_1d(g["misc_x"], g["misc_y"]), g["misc_ev"], rtol=1e-12)
    x = g["misc_x"]
    np.testing.assert_allclose((x - x.mean()) / (x.std() + 1e-8), g["misc_center"], rtol=1e-12)
    # discount_cumsum([1,1,1], .5) = [1.75, 1.5, 1]
    np.testing.assert_allclose(S.discount_cumsum(np.ones(3), 0.5), [1.75, 1.5, 1.0])


def test_lane_rollout_reproduces_reference_rollout_semantics(golden):
    """rollout_lanes with reset_states replay == per-path reference rollout on PointEnv, incl. auto-reset."""
    g = golden
    env = E.make("point")
    dims = P.Dims(2, (4, 4), 2)
    theta = np.zeros(dims.P)
    theta[-2:] = -30.0                # sigma ~ 1e-13: action == eps*sigma + mean(=0) ~ 0
    T = 12
    acts = g["pt_done_actions"]
    # use eps to inject actions exactly: set log_std = 0 and mean = 0 -> action = eps
    theta[-2:] = 0.0
    eps = np.zeros((T, 2, 1))
    eps[:10, :, 0] = acts
    rs = np.zeros((T + 1, 2, 1))
    rs[:, 0, 0], rs[:, 1, 0] = 0.05, -0.03
    traj = S.rollout_lanes(env, theta, dims, 1, T, 100, eps, None, reset_states=rs)
    L = int(g["pt_done_len"])
    assert traj["flags"][L - 1, 0] == (S.FLAG_DONE | S.FLAG_END)
    assert np.array_equal(traj["obs"][:, :L, 0].T, g["pt_done_obs"])
    assert np.array_equal(traj["rew"][:L, 0], g["pt_done_rew"])
    assert traj["tstep"][L, 0] == 0 and np.array_equal(traj["obs"][:, L, 0], [0.05, -0.03])


def test_swimmer_learning_curve_fixture_is_sane():
    """tests/golden/oracle_swimmer_trpo_curve.json (make_swimmer_curve.py): the oracle's TRPO run on Swimmer that the GPU
    learning check compares against -- every step accepted within the trust region, return improving."""
    import json
    import os
    d = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_swimmer_trpo_curve.json")))
    cfg, curve = d["config"], d["curve"]
    assert (cfg["lanes"], cfg["horizon"], cfg["cg_iters"], cfg["step_size"]) == (1024, 500, 10, 0.01)
    assert [r["itr"] for r in curve] == list(range(40))
    assert all(r["NumTrajs"] == cfg["lanes"] for r in curve)                  # Swimmer never terminates early
    assert all((not r["rejected"]) and 0 < r["MeanKL"] <= cfg["step_size"] for r in curve)
    assert all(r["LossAfter"] < r["LossBefore"] for r in curve)
    ret = np.array([r["AverageReturn"] for r in curve])
    assert ret[0] < 0 < ret[10] < ret[20] < ret[39] and ret[39] > 30


def test_hopper_learning_curve_fixture_is_sane():
    """tests/golden/oracle_hopper_trpo_curve.json: the same oracle TRPO run on Hopper with the cfg4 net (64,64)."""
    import json
    import os
    d = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_hopper_trpo_curve.json")))
    cfg, curve = d["config"], d["curve"]
    assert (cfg["env"], cfg["hidden"], cfg["lanes"], cfg["horizon"]) == ("hopper", [64, 64], 1024, 500)
    assert [r["itr"] for r in curve] == list(range(40))
    assert all((not r["rejected"]) and 0 < r["MeanKL"] <= cfg["step_size"] for r in curve)
    ret = np.array([r["AverageReturn"] for r in curve])
    ntraj = np.array([r["NumTrajs"] for r in curve])
    assert np.all(np.diff(ret) > 0) and ret[-1] > 200          # monotone improvement: the hopper stays up longer ...
    assert ntraj[0] > 10 * ntraj[-1] >= cfg["lanes"]           # ... so the same 512 000 samples hold far fewer paths


def test_planar_models_match_reference_mujoco_xml():
    """oracle/planar.py's Swimmer / Hopper model constants re-derived from the reference's own model files
    (tests/golden/reference_mujoco_models.json <- vendor/mujoco_models/*.xml, make_mujoco_model_golden.py)."""
    import json
    import os
    from oracle import planar as PL
    ref = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_mujoco_models.json")))

    def check(name, model, plane, hinge_names, geom_density):
        x = ref[name]
        ix, iy = plane                                           # which world axes span the model's plane
        bodies = x["bodies"]
        assert len(bodies) == model.n
        assert model.dt == x["option"]["timestep"][0]
        assert model.rk4 == (x["option"]["integrator"] == "RK4")
        assert model.frame_skip == int(x["custom"].get("frame_skip", [1])[0])
        assert model.density == x["option"].get("density", [0.0])[0]
        assert model.viscosity == x["option"].get("viscosity", [0.0])[0]
        global_coords = x["compiler"]["coordinate"] == "global"
        jdef = x["default"].get("joint", {})
        hinge_world = []                                         # world position of each body's hinge (reference pose)
        origin_world = []
        for b in bodies:
            hinge = [j for j in b["joints"] if j["name"] == hinge_names[len(hinge_world)]][0]
            if global_coords:
                hw, ow = np.array(hinge["pos"]), np.array(b["pos"])
            else:                                                # local: positions relative to the parent body frame
                parent = [i for i, p in enumerate(bodies) if p["name"] == b["parent"]]
                base = origin_world[parent[0]] if parent else np.zeros(3)
                ow = base + np.array(b["pos"])
                hw = ow + np.array(hinge["pos"])
            hinge_world.append(hw), origin_world.append(ow)
        for i, b in enumerate(bodies):
            g = b["geoms"][0]
            p0, p1 = np.array(g["fromto"][:3]), np.array(g["fromto"][3:])
            if not global_coords:
                p0, p1 = p0 + origin_world[i], p1 + origin_world[i]
            r, L = g["size"][0], np.linalg.norm(p1 - p0)
            mass, Ip, Ia = PL.capsule(r, L, g.get("density", [geom_density])[0])
            np.testing.assert_allclose([model.mass[i], model.Ip[i], model.Ia[i]], [mass, Ip, Ia], rtol=1e-12)
            com = 0.5 * (p0 + p1) - hinge_world[i]
            np.testing.assert_allclose(model.c[i], (com[ix], com[iy]), atol=1e-12)
            axis = (p1 - p0) / L
            np.testing.assert_allclose(np.abs(model.long_axis[i]), np.abs((axis[ix], axis[iy])), atol=1e-12)
            parent = [k for k, p in enumerate(bodies) if p["name"] == b["parent"]]
            anchor = hinge_world[i] - (hinge_world[parent[0]] if parent else hinge_world[i])
            np.testing.assert_allclose(model.a[i], (anchor[ix], anchor[iy]), atol=1e-12)
            bo = origin_world[i] - hinge_world[i]
            np.testing.assert_allclose(model.bo[i], (bo[ix], bo[iy]), atol=1e-12)
            hinge = [j for j in b["joints"] if j["name"] == hinge_names[i]][0]
            limited = hinge.get("limited", jdef.get("limited", "false")) == "true" and "range" in hinge
            if limited:
                np.testing.assert_allclose(model.limits[i], np.deg2rad(hinge["range"]), rtol=1e-12)
            else:
                assert model.limits[i] is None
        # degrees of freedom in qpos order = joint order of the XML
        joints = [j for b in bodies for j in b["joints"]]
        assert len(joints) == len(model.armature) == len(model.damping) == len(model.q0)
        for k, j in enumerate(joints):
            assert model.armature[k] == j.get("armature", jdef.get("armature", [0.0]))[0], j["name"]
            assert model.damping[k] == j.get("damping", jdef.get("damping", [0.0]))[0], j["name"]
            assert model.q0[k] == j.get("ref", [0.0])[0], j["name"]
        act = [a["joint"] for a in x["actuators"]]
        assert [hinge_names[i] for i in model.act] == act
        assert all(a["ctrlrange"] == [-model.ctrl_lim, model.ctrl_lim] for a in x["actuators"])
        return bodies, hinge_world

    sw = PL.swimmer_model()
    check("swimmer", sw, (0, 1), ["rot", "rot2", "rot3"], 1000.0)
    hp = PL.hopper_model()
    bodies, hinge_world = check("hopper", hp, (0, 2), ["rooty", "thigh_joint", "leg_joint", "foot_joint"], 1000.0)
    assert hp.gravity == (0.0, -9.81)                            # MuJoCo's default gravity, along -z
    gdef = ref["hopper"]["default"]["geom"]
    foot = bodies[3]["geoms"][0]
    assert hp.mu == foot["friction"][0] and hp.margin == gdef["margin"][0]
    assert list(hp.con_solref) == gdef["solref"] and list(hp.con_solimp) == gdef["solimp"]
    ends = [np.array(foot["fromto"][:3]) - hinge_world[3], np.array(foot["fromto"][3:]) - hinge_world[3]]
    for c, e in zip(hp.contacts, ends):                          # the two end spheres of the foot capsule
        assert c["body"] == 3 and c["r"] == foot["size"][0]
        np.testing.assert_allclose(c["e"], (e[0], e[2]), atol=1e-12)


def test_cartpole_model_matches_reference_box2d_template():
    """oracle/envs.py::CartPoleEnv's reduced-coordinate constants re-derived from the reference's Box2D model template and
    env class (tests/golden/reference_cartpole_model.json <- models/cartpole.xml.mako, cartpole_env.py)."""
    import json
    import os
    from oracle import envs as E
    r = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_cartpole_model.json")))
    e = E.CartPoleEnv
    assert r["densities"] == [1.0, 1.0]
    np.testing.assert_allclose(e.M, r["densities"][0] * r["cart_width"] * r["cart_height"], rtol=1e-12)     # box area x density
    np.testing.assert_allclose(e.m, r["densities"][1] * r["pole_width"] * r["pole_height"], rtol=1e-12)
    # the pole rectangle runs from the hinge (0, 0) to (0, pole_height): COM half-way up, hinge on the cart's top edge
    assert r["pole_vertices_expr"] == "compute_rect_vertices((0, 0), (0, pole_height), pole_width/2)"
    assert r["pole_anchor_is_cart_top"]
    np.testing.assert_allclose(e.l, r["pole_height"] / 2, rtol=1e-12)
    np.testing.assert_allclose(e.I, e.m * (r["pole_width"] ** 2 + r["pole_height"] ** 2) / 12.0, rtol=1e-12)
    assert e.dt_ == r["timestep"]
    assert (e.lb[0], e.ub[0]) == tuple(r["ctrllimit"])
    assert e.bounds == (r["max_cart_pos"], r["max_cart_speed"], r["max_pole_angle"], r["max_pole_speed"])
    assert e.reset_range == r["reset_range"]


def test_double_pendulum_model_matches_reference_box2d_template():
    """oracle/envs.py::DoublePendulumEnv's constants re-derived from the reference's template and env class
    (tests/golden/reference_double_pendulum_model.json <- models/double_pendulum.xml.mako, double_pendulum_env.py), plus
    two physical sanity checks of the restated dynamics: energy is conserved without torque (to the integrator's order)
    and the hanging equilibrium is a fixed point."""
    import json
    import os
    from oracle import envs as E
    r = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden",
                                    "reference_double_pendulum_model.json")))
    e = E.DoublePendulumEnv
    L = float(r["link_len_default"])
    assert r["densities"] == [5.0, 5.0] and e.L == L
    np.testing.assert_allclose(e.m, r["densities"][0] * r["link_width"] * L, rtol=1e-12)        # rectangle area x density
    assert all(v.replace(" ", "") == "compute_rect_vertices([0,0],[0,-link_len],link_width/2)" for v in r["vertices_exprs"])
    np.testing.assert_allclose(e.lc, L / 2, rtol=1e-12)                                         # COM half-way down the rod
    np.testing.assert_allclose(e.I, e.m * (r["link_width"] ** 2 + L ** 2) / 12.0, rtol=1e-12)
    assert [tuple(j) for j in r["joints"]] == [("revolute", "link_joint_1", "track", "link1", "0,0"),
                                               ("revolute", "link_joint_2", "link1", "link2", "0,${-link_len}")]
    assert r["controls"] == [["torque", "link_joint_2", e.lb[0], e.ub[0]]]
    assert [tuple(s) for s in r["states"]] == [("apos", "link1", "sin"), ("apos", "link1", "cos"), ("avel", "link1", ""),
                                               ("apos", "link2", "sin"), ("apos", "link2", "cos"), ("avel", "link2", "")]
    assert e.dt_ == r["timestep"] and e.frame_skip == r["frame_skip_default"] and r["never_done"]
    assert r["reset_stds"] == [0.1, 0.1, 0.01, 0.01]
    assert r["tip_formula"] == ["cur_center_pos[0] - self.link_len*np.sin(cur_angle),",
                                "cur_center_pos[1] - self.link_len*np.cos(cur_angle)"]
    env = E.DoublePendulumEnv(np.float64)
    raw = np.array([[3.0], [-2.0], [1.0], [0.5]])
    s0 = env.reset(raw)
    np.testing.assert_allclose(s0[:, 0], [0.3, -0.2, 0.01, 0.005], rtol=1e-12)
    np.testing.assert_allclose(env.obs(s0)[:, 0], [np.sin(0.3), np.cos(0.3), 0.01, np.sin(-0.2), np.cos(-0.2), 0.005])

    def energy(s):
        th1, th2, w1, w2 = s[:, 0]
        m, lc, Lk, I, g = env.m, env.lc, env.L, env.I, env.g
        kin = 0.5 * (I + m * lc * lc + m * Lk * Lk) * w1 * w1 + 0.5 * (I + m * lc * lc) * w2 * w2 \
            + m * Lk * lc * np.cos(th1 - th2) * w1 * w2
        pot = -m * g * lc * np.cos(th1) - m * g * (Lk * np.cos(th1) + lc * np.cos(th2))
        return kin + pot
    # the equations of motion conserve the Lagrangian's energy: the drift of the semi-implicit Euler scheme over 4 s of
    # free swinging shrinks with the step (first order), 100x from dt = 1e-2 to 1e-4
    e_min = -env.m * env.g * (2 * env.lc + env.L)
    drift = {}
    for h, n in ((1e-2, 200), (1e-4, 20000)):
        env.dt_, s = h, s0
        for _ in range(n):
            s, rew, done = env.step(s, np.zeros((1, 1)))
        drift[h] = abs(energy(s) - energy(s0)) / (energy(s0) - e_min)
        assert not done.any()
    env.dt_ = E.DoublePendulumEnv.dt_
    assert drift[1e-4] < 5e-3 and drift[1e-4] < 0.05 * drift[1e-2], drift
    rest = np.zeros((4, 1))
    s1, rew, _ = env.step(rest, np.zeros((1, 1)))
    np.testing.assert_allclose(s1, 0.0, atol=1e-15)
    np.testing.assert_allclose(rew, -4.0 * L)                                       # hanging tip (0,-2L) vs target (0,2L)
    # tip formula of the reference (its x sign included): th1 = pi/2, th2 = pi/2 -> link2 origin (L, 0), tip (L - L, 0 - 0)
    s2 = np.array([[np.pi / 2], [np.pi / 2], [0.0], [0.0]])
    tx = L * np.sin(s2[0]) - L * np.sin(s2[1])
    assert abs(tx[0]) < 1e-15

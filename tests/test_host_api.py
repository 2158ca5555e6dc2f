"""CPU: host-side pieces of the plugin API that need no GPU (reference known answers where they exist)."""
import pickle

import numpy as np
import pytest


def test_truncate_paths_reference_known_answer(golden):
    """tests/test_sampler.py:4-32 of the reference, plus the golden lengths produced by the reference itself."""
    from rllab_b200.sampler.parallel_sampler import truncate_paths
    mk = lambda l: dict(observations=np.zeros((l, 1)), actions=np.zeros((l, 1)), rewards=np.zeros(l), env_infos=dict(),
                        agent_infos=dict(lala=np.zeros(l)))
    paths = [mk(100), mk(50)]
    truncated = truncate_paths(paths, 130)
    assert len(truncated) == 2 and len(truncated[-1]["observations"]) == 30 and len(truncated[0]["observations"]) == 100
    assert len(paths) == 2 and len(paths[-1]["observations"]) == 50          # input not modified
    assert len(truncated[-1]["agent_infos"]["lala"]) == 30
    for ms in (130, 150, 1, 249, 250, 400):
        got = [len(p["rewards"]) for p in truncate_paths([mk(int(l)) for l in golden["tr_lens"]], ms)]
        assert got == list(golden["tr_%d" % ms])


def test_box_space_reference_semantics():
    """tests/test_spaces.py:20-27 (Box flatten / unflatten round trips)."""
    from rllab_b200.spaces import Box
    b = Box(-1.0, 1.0, (3, 4))
    x = b.sample()
    assert b.contains(x) and b.flat_dim == 12 and b.shape == (3, 4)
    np.testing.assert_array_equal(b.unflatten(b.flatten(x)), x)
    xs = np.stack([b.sample() for _ in range(5)])
    np.testing.assert_array_equal(b.unflatten_n(b.flatten_n(xs)), xs)
    assert Box(np.array([-1., -2.]), np.array([2., 4.])) == Box(np.array([-1., -2.]), np.array([2., 4.]))
    assert not b.contains(np.full((3, 4), 2.0))


def test_linear_feature_baseline_host_api_matches_reference_golden(golden):
    """fit(paths)/predict(path) on host path dicts == the reference's own LinearFeatureBaseline (golden ps_*_fit)."""
    from oracle import sampler as S
    from rllab_b200.baselines.linear_feature_baseline import LinearFeatureBaseline
    traj = {k[len("ps_in_"):]: v for k, v in golden.items() if k.startswith("ps_in_") and k != "ps_in_coeffs_prev"}
    paths = S.lanes_to_paths(traj)
    out = S.process_samples_lanes(traj, None, 0.99, 1.0)
    for p in paths:
        L = len(p["rewards"])
        p["returns"] = out["ret"][p["_t0"]:p["_t0"] + L, p["_lane"]]
    bl = LinearFeatureBaseline()
    assert np.array_equal(bl.predict(paths[0]), np.zeros(len(paths[0]["rewards"])))     # :41-42
    bl.fit(paths)
    pred = np.concatenate([bl.predict(p) for p in paths])
    bl_ref = LinearFeatureBaseline()
    bl_ref.set_param_values(golden["ps_a_fit"])
    np.testing.assert_allclose(pred, np.concatenate([bl_ref.predict(p) for p in paths]), rtol=1e-6, atol=1e-8)
    bl2 = pickle.loads(pickle.dumps(bl))
    np.testing.assert_array_equal(bl2.get_param_values(), bl.get_param_values())


def test_diagonal_gaussian_host_api_matches_reference_golden(golden):
    from rllab_b200.distributions.diagonal_gaussian import DiagonalGaussian
    g = golden
    d = DiagonalGaussian(2)
    np.testing.assert_allclose(d.kl(dict(mean=g["dg_om"], log_std=g["dg_ol"]), dict(mean=g["dg_nm"], log_std=g["dg_nl"])),
                               g["dg_kl"], rtol=1e-13)
    np.testing.assert_allclose(d.log_likelihood(g["dg_xs"], dict(mean=g["dg_nm"], log_std=g["dg_nl"])), g["dg_ll"], rtol=1e-13)
    np.testing.assert_allclose(d.entropy(dict(log_std=g["dg_nl"])), g["dg_ent"], rtol=1e-13)
    assert d.dist_info_keys == ["mean", "log_std"] and d.dim == 2


def test_policy_construction_and_errors_without_gpu():
    from rllab_b200 import _lib
    from rllab_b200.envs.box2d.cartpole_env import CartpoleEnv
    from rllab_b200.envs.normalized_env import normalize
    from rllab_b200.policies.gaussian_mlp_policy import GaussianMLPPolicy
    env = normalize(CartpoleEnv())
    assert env.action_space.bounds[0][0] == -1.0 and env.observation_space.flat_dim == 4
    pol = GaussianMLPPolicy(env.spec, hidden_sizes=(32, 32), seed=4)
    flat = pol.get_param_values()
    assert flat.shape == (1250,) and np.all(flat[-1:] == 0.0)                    # log(init_std=1)
    W0 = pol.flat_to_params(flat)[0]
    assert W0.shape == (4, 32) and np.abs(W0).max() <= np.sqrt(6.0 / 36) + 1e-12   # GlorotUniform bound
    pol.set_param_values(flat * 2)
    np.testing.assert_array_equal(pol.get_param_values(), flat * 2)
    pol2 = pickle.loads(pickle.dumps(pol))
    np.testing.assert_array_equal(pol2.get_param_values(), flat * 2)
    with pytest.raises(NotImplementedError):
        GaussianMLPPolicy(env.spec, adaptive_std=True)
    with pytest.raises(_lib.B200RLError):
        GaussianMLPPolicy(env.spec, hidden_sizes=(16, 16))
    with pytest.raises(NotImplementedError):
        normalize(CartpoleEnv(), normalize_obs=True)
    from rllab_b200.optimizers.conjugate_gradient_optimizer import ConjugateGradientOptimizer
    ConjugateGradientOptimizer(subsample_factor=0.5)            # built in round 2 (tile-granular sub-sampling)
    with pytest.raises(ValueError):
        ConjugateGradientOptimizer(subsample_factor=0.0)
    with pytest.raises(TypeError):
        ConjugateGradientOptimizer(hvp_approach="finite-difference")


def test_logger_tabular_and_snapshot(tmp_path):
    from rllab_b200.misc import logger
    logger.set_quiet(True)
    logger.set_snapshot_dir(str(tmp_path))
    logger.set_snapshot_mode("last")
    with logger.prefix("itr #0 | "):
        logger.record_tabular("AverageReturn", 1.5)
        logger.record_tabular("NumTrajs", 7)
        logger.dump_tabular()
    assert logger.get_last_table() == {"AverageReturn": 1.5, "NumTrajs": 7}
    # deferred values (device readbacks) are resolved when the table is dumped, not when they are recorded
    box = {"v": 1.0}
    logger.record_tabular("LossBefore", lambda: box["v"])
    box["v"] = 2.0
    logger.dump_tabular()
    assert logger.get_last_table() == {"LossBefore": 2.0}
    logger.save_itr_params(0, dict(itr=0, x=np.arange(3)))
    d = pickle.load(open(str(tmp_path / "params.pkl"), "rb"))
    assert d["itr"] == 0
    logger.set_snapshot_mode("none")
    logger.set_snapshot_dir(None)


def test_plugin_api_surface_matches_reference():
    """Constructor arguments (names, literal defaults) and public members of every mirrored class against the
    reference's own source (tests/golden/reference_api.json, extracted with ast by tests/golden/make_api_golden.py).
    Only Theano-symbolic members are exempt: this implementation has no symbolic graph (the C ABI replaces it)."""
    import importlib
    import inspect
    import json
    import os
    api = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_api.json")))
    exempt = {
        "Box": {"new_tensor_variable"},
        "DiagonalGaussian": {"kl_sym", "likelihood_ratio_sym", "log_likelihood_sym", "entropy_sym"},
        "GaussianMLPPolicy": {"dist_info_sym", "get_reparam_action_sym"},
        "FirstOrderOptimizer": {"optimize_gen"},      # generator form of the mini-batch loop; VPG uses optimize()
        # simulator internals (Box2D generators, MuJoCo model accessors): the dynamics live in the CUDA kernels
        "CartpoleEnv": {"compute_reward", "is_current_done"},
        "SwimmerEnv": {"get_current_obs", "get_ori"},
        "HopperEnv": {"get_current_obs"},
    }
    tabular = api.pop("__tabular__")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for rel, d in tabular.items():                     # same logger.record_tabular keys as the reference records
        src = "".join(open(os.path.join(root, m)).read() for m in d["mirrors"])
        assert d["keys"], rel
        for key in d["keys"]:
            assert ('"%s"' % key in src) or ("'%s'" % key in src) or (key.endswith("ForwardProgress") and "ForwardProgress" in src), (rel, key)
    assert len(api) >= 19
    for name, d in sorted(api.items()):
        mod, cls = d["mirror"].rsplit(".", 1)
        C = getattr(importlib.import_module(mod), cls)
        params = {}
        if inspect.isfunction(C):                      # GymEnv is a factory returning the device-backed PendulumEnv
            params = dict(inspect.signature(C).parameters)
            C = getattr(importlib.import_module(mod), "PendulumEnv")
        else:
            for k in C.__mro__:                        # the reference forwards **kwargs up its class chain as well
                if "__init__" in k.__dict__:
                    for p in inspect.signature(k.__dict__["__init__"]).parameters.values():
                        params.setdefault(p.name, p)
        for a in (d["init"] or {"args": []})["args"]:
            assert a["name"] in params, "%s.__init__ lacks the reference argument %r" % (name, a["name"])
            if a["default"] and "literal" in a["default"]:
                mine = params[a["name"]].default
                assert mine is not inspect.Parameter.empty, (name, a["name"])
                mine = list(mine) if isinstance(mine, tuple) else mine
                assert mine == a["default"]["literal"], (name, a["name"], mine, a["default"]["literal"])
        for m in d["methods"] + d["properties"]:
            if m in exempt.get(name, ()):
                assert not hasattr(C, m)               # keep the exemption list honest
                continue
            assert hasattr(C, m), "%s lacks the reference member %r" % (name, m)


def test_env_constructors_accept_reference_arguments_and_reject_other_values():
    """hopper_env.py:27-35, swimmer_env.py:17-23, cartpole_env.py:13-24: the reference defaults are what the CUDA
    dynamics implement; other values fail loudly (no silent fallback), unknown names are TypeErrors."""
    from rllab_b200.envs.box2d.cartpole_env import CartpoleEnv
    from rllab_b200.envs.mujoco.hopper_env import HopperEnv
    from rllab_b200.envs.mujoco.swimmer_env import SwimmerEnv
    assert SwimmerEnv(ctrl_cost_coeff=1e-2).ctrl_cost_coeff == 1e-2
    h = HopperEnv(alive_coeff=1, ctrl_cost_coeff=0.01, action_noise=0.0)
    assert (h.alive_coeff, h.ctrl_cost_coeff) == (1, 0.01)
    assert CartpoleEnv(frame_skip=1).max_cart_pos == 2.4
    for bad in (lambda: SwimmerEnv(ctrl_cost_coeff=0.1), lambda: HopperEnv(alive_coeff=2), lambda: CartpoleEnv(obs_noise=0.1)):
        with pytest.raises(NotImplementedError):
            bad()
    with pytest.raises(TypeError):
        HopperEnv(bogus=1)

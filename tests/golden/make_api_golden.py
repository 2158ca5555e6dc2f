"""Generate tests/golden/reference_api.json: the constructor signatures and public method names of the reference classes
that rllab_b200 mirrors, extracted from the reference SOURCE with `ast` (nothing is imported, so Theano is not needed).

Run:  python tests/golden/make_api_golden.py        (needs /root/reference; the tests only read the committed JSON)
"""
import ast
import json
import os

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_api.json")

# reference file -> {class name: rllab_b200 dotted path of the mirror}
CLASSES = {
    "rllab/algos/batch_polopt.py": {"BatchPolopt": "rllab_b200.algos.batch_polopt.BatchPolopt"},
    "rllab/algos/npo.py": {"NPO": "rllab_b200.algos.npo.NPO"},
    "rllab/algos/trpo.py": {"TRPO": "rllab_b200.algos.trpo.TRPO"},
    "rllab/algos/vpg.py": {"VPG": "rllab_b200.algos.vpg.VPG"},
    "rllab/policies/gaussian_mlp_policy.py": {"GaussianMLPPolicy": "rllab_b200.policies.gaussian_mlp_policy.GaussianMLPPolicy"},
    "rllab/optimizers/conjugate_gradient_optimizer.py": {
        "ConjugateGradientOptimizer": "rllab_b200.optimizers.conjugate_gradient_optimizer.ConjugateGradientOptimizer"},
    "rllab/optimizers/first_order_optimizer.py": {
        "FirstOrderOptimizer": "rllab_b200.optimizers.first_order_optimizer.FirstOrderOptimizer"},
    "rllab/baselines/linear_feature_baseline.py": {
        "LinearFeatureBaseline": "rllab_b200.baselines.linear_feature_baseline.LinearFeatureBaseline"},
    "rllab/baselines/zero_baseline.py": {"ZeroBaseline": "rllab_b200.baselines.zero_baseline.ZeroBaseline"},
    "rllab/envs/normalized_env.py": {"NormalizedEnv": "rllab_b200.envs.normalized_env.NormalizedEnv"},
    "rllab/envs/base.py": {"Env": "rllab_b200.envs.base.Env"},
    "rllab/spaces/box.py": {"Box": "rllab_b200.spaces.box.Box"},
    "rllab/distributions/diagonal_gaussian.py": {
        "DiagonalGaussian": "rllab_b200.distributions.diagonal_gaussian.DiagonalGaussian"},
    "rllab/sampler/base.py": {"Sampler": "rllab_b200.sampler.base.Sampler"},
    "rllab/envs/box2d/cartpole_env.py": {"CartpoleEnv": "rllab_b200.envs.box2d.cartpole_env.CartpoleEnv"},
    "rllab/envs/mujoco/swimmer_env.py": {"SwimmerEnv": "rllab_b200.envs.mujoco.swimmer_env.SwimmerEnv"},
    "rllab/envs/mujoco/hopper_env.py": {"HopperEnv": "rllab_b200.envs.mujoco.hopper_env.HopperEnv"},
    "examples/point_env.py": {"PointEnv": "rllab_b200.envs.point_env.PointEnv"},
    "rllab/envs/gym_env.py": {"GymEnv": "rllab_b200.envs.gym_env.GymEnv"},
}


# reference file -> mirror source files that must record the same logger.record_tabular keys
TABULAR = {
    "rllab/sampler/base.py": ["rllab_b200/sampler/lane_sampler.py"],
    "rllab/algos/vpg.py": ["rllab_b200/algos/vpg.py"],
    "rllab/algos/npo.py": ["rllab_b200/algos/npo.py"],
    "rllab/policies/gaussian_mlp_policy.py": ["rllab_b200/algos/batch_polopt.py"],
    "rllab/envs/mujoco/swimmer_env.py": ["rllab_b200/envs/mujoco/swimmer_env.py"],
    "rllab/envs/mujoco/hopper_env.py": ["rllab_b200/envs/mujoco/hopper_env.py"],
}


def tabular_keys(tree):
    keys = []
    for n in ast.walk(tree):
        if (isinstance(n, ast.Call) and isinstance(n.func, ast.Attribute) and n.func.attr == "record_tabular" and n.args
                and isinstance(n.args[0], ast.Constant) and n.args[0].value not in keys):
            keys.append(n.args[0].value)
    return keys


def literal(node):
    try:
        return {"literal": ast.literal_eval(node)}
    except Exception:
        return {"source": ast.unparse(node)}


def describe(cls):
    out = {"init": None, "methods": [], "properties": []}
    for item in cls.body:
        if not isinstance(item, ast.FunctionDef):
            continue
        decorators = [ast.unparse(d) for d in item.decorator_list]
        if item.name == "__init__":
            a = item.args
            names = [x.arg for x in a.args][1:]
            defaults = [None] * (len(names) - len(a.defaults)) + [literal(d) for d in a.defaults]
            out["init"] = {"args": [{"name": n, "default": d} for n, d in zip(names, defaults)],
                           "kwargs": a.kwarg is not None}
        elif not item.name.startswith("_"):
            (out["properties"] if "property" in decorators else out["methods"]).append(item.name)
    return out


def main():
    api = {}
    for rel, classes in CLASSES.items():
        tree = ast.parse(open(os.path.join(REF, rel)).read())
        for node in tree.body:
            if isinstance(node, ast.ClassDef) and node.name in classes:
                d = describe(node)
                d["mirror"] = classes[node.name]
                d["reference_file"] = rel
                api[node.name] = d
    missing = {c for cl in CLASSES.values() for c in cl} - set(api)
    assert not missing, missing
    api["__tabular__"] = {rel: {"keys": tabular_keys(ast.parse(open(os.path.join(REF, rel)).read())), "mirrors": mirrors}
                          for rel, mirrors in TABULAR.items()}
    with open(OUT, "w") as f:
        json.dump(api, f, indent=1, sort_keys=True, default=str)
    print("wrote", OUT, len(api) - 1, "classes")


if __name__ == "__main__":
    main()

"""Generate tests/golden/reference_cartpole_model.json: the constants of the reference's Box2D CartPole, read from
rllab/envs/box2d/models/cartpole.xml.mako (the Python block of the template is executed on its own, the XML attributes are
read with regular expressions) and from rllab/envs/box2d/cartpole_env.py (ast).  tests/test_oracle_golden.py re-derives the
reduced-coordinate constants of oracle/envs.py::CartPoleEnv (cart / pole mass, pole half length and inertia, time step,
force limit, termination bounds, reset range) from it.  Box2D's solver itself is third-party: parity unpinned.

Run:  python tests/golden/make_box2d_model_golden.py        (needs /root/reference)
"""
import ast
import json
import os
import re

REF = "/root/reference/rllab/envs/box2d"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_cartpole_model.json")


def double_pendulum():
    """tests/golden/reference_double_pendulum_model.json: constants of models/double_pendulum.xml.mako and
    double_pendulum_env.py (link length / width / density, joints, control limit, time step, frame_skip default, reset
    standard deviations)."""
    mako = open(os.path.join(REF, "models", "double_pendulum.xml.mako")).read()
    out = dict(link_width=float(re.search(r"link_width = ([0-9.]+)", mako).group(1)),
               timestep=float(re.search(r'<world timestep="([^"]+)"', mako).group(1)),
               densities=[float(x) for x in re.findall(r'density="([^"]+)"', mako)],
               vertices_exprs=re.findall(r'vertices="\$\{([^}]+)\}"', mako),
               joints=re.findall(r'<joint type="(\w+)" name="(\w+)" bodyA="(\w+)" bodyB="(\w+)" anchor="([^"]+)"', mako),
               states=re.findall(r'<state type="(\w+)" body="(\w+)"(?: transform="(\w+)")?', mako))
    ctrl = [m for m in re.finditer(r'^\s*<control type="(\w+)" joint="(\w+)" ctrllimit="([^"]+)"', mako, re.M)]
    out["controls"] = [[m.group(1), m.group(2)] + [float(x) for x in m.group(3).split(",")] for m in ctrl]
    src = open(os.path.join(REF, "double_pendulum_env.py")).read()
    out["frame_skip_default"] = int(re.search(r'kwargs.get\("frame_skip", (\d+)\)', src).group(1))
    out["link_len_default"] = int(re.search(r"self.link_len = (\d+)\n", src).group(1))
    out["reset_stds"] = [float(x) for x in re.search(r"stds = np.array\(\[([^\]]+)\]\)", src).group(1).split(",")]
    out["tip_formula"] = [l.strip() for l in src.splitlines() if "link_len*np." in l]
    out["never_done"] = "def is_current_done(self):\n        return False" in src
    path = os.path.join(os.path.dirname(OUT), "reference_double_pendulum_model.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote", path, out)


def main():
    double_pendulum()
    mako = open(os.path.join(REF, "models", "cartpole.xml.mako")).read()
    block = mako[mako.index("<%") + 2:mako.index("%>")]
    block = "\n".join(l for l in block.splitlines() if "import compute_rect_vertices" not in l)
    env = {"opts": {}}
    exec("\n".join(l[4:] if l.startswith("    ") else l for l in block.splitlines()), env)
    out = {k: env[k] for k in ("cart_width", "cart_height", "pole_width", "pole_height", "cart_friction", "pole_friction")}
    out["timestep"] = float(re.search(r'<world timestep="([^"]+)"', mako).group(1))
    out["densities"] = [float(x) for x in re.findall(r'density="([^"]+)"', mako)]
    lo, hi = re.search(r'ctrllimit="([^"]+)"', mako).group(1).split(",")
    out["ctrllimit"] = [float(lo), float(hi)]
    out["pole_anchor_is_cart_top"] = 'anchor="0,${cart_height}"' in mako
    out["pole_vertices_expr"] = re.search(r'vertices="\$\{([^}]+)\}"', mako).group(1)
    tree = ast.parse(open(os.path.join(REF, "cartpole_env.py")).read())
    for n in ast.walk(tree):
        if isinstance(n, ast.Assign) and isinstance(n.targets[0], ast.Attribute) and isinstance(n.value, ast.Constant):
            out[n.targets[0].attr] = n.value.value
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote", OUT, out)


if __name__ == "__main__":
    main()

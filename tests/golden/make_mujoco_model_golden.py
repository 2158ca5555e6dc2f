"""Generate tests/golden/reference_mujoco_models.json: the physical parameters of the reference's Swimmer and Hopper models,
read from vendor/mujoco_models/{swimmer,hopper}.xml with xml.etree (nothing is executed).  tests/test_oracle_golden.py
re-derives the constants of oracle/planar.py (capsule masses and inertias, hinge anchors, centres of mass, joint limits,
damping / armature, actuator limits, contact geometry and solver parameters, time step / integrator / frame skip) from
this fixture, which pins the restated MODEL to the reference's files (the dynamics ALGORITHM remains a restatement of the
published MuJoCo pipeline: "parity unpinned", oracle/__init__.py).

Run:  python tests/golden/make_mujoco_model_golden.py        (needs /root/reference)
"""
import json
import os
import xml.etree.ElementTree as ET

REF = "/root/reference/vendor/mujoco_models"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_mujoco_models.json")


def nums(s):
    # hopper.xml writes the foot body position as "0.13/2 0 0.1": plain arithmetic in an attribute
    return [float(eval(tok, {"__builtins__": {}}, {})) for tok in s.split()]


def attrs(e, numeric):
    d = {}
    for k, v in e.attrib.items():
        d[k] = nums(v) if k in numeric else v
    return d


NUMERIC = {"pos", "axis", "range", "fromto", "size", "density", "friction", "damping", "armature", "ref", "stiffness",
           "ctrlrange", "margin", "solref", "solimp", "timestep", "viscosity", "data"}


def bodies(e, parent, out):
    for b in e.findall("body"):
        out.append(dict(name=b.get("name"), parent=parent, pos=nums(b.get("pos")),
                        joints=[attrs(j, NUMERIC) for j in b.findall("joint")],
                        geoms=[attrs(g, NUMERIC) for g in b.findall("geom")]))
        bodies(b, b.get("name"), out)


def model(path):
    root = ET.parse(path).getroot()
    m = dict(compiler=dict(root.find("compiler").attrib), option=attrs(root.find("option"), NUMERIC))
    default = root.find("default")
    m["default"] = {c.tag: attrs(c, NUMERIC) for c in default} if default is not None else {}
    custom = root.find("custom")
    m["custom"] = {n.get("name"): nums(n.get("data")) for n in custom.findall("numeric")} if custom is not None else {}
    m["bodies"] = []
    bodies(root.find("worldbody"), None, m["bodies"])
    m["actuators"] = [attrs(a, NUMERIC) for a in root.find("actuator").findall("motor")]
    return m


if __name__ == "__main__":
    out = {name: model(os.path.join(REF, name + ".xml")) for name in ("swimmer", "hopper")}
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote", OUT)

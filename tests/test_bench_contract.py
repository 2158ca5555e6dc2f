"""CPU: the CPU leg of bench.py (the `cpu_baseline` object / the `--impl reference` arm) runs in a clean interpreter and
returns the keys the bench contract names.  (The GPU arm of the contract is exercised on the GPU box by bench.py itself.)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpu_baseline_leg_contract():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--cpu-arm-json", "--cpu-seconds", "3"],
                         capture_output=True, text=True, timeout=300, env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    assert out.returncode == 0, out.stderr[-2000:]
    info = json.loads(out.stdout.strip().splitlines()[-1])
    assert set(info) >= {"value", "unit", "cores", "kind", "sample"}
    assert info["unit"] == "env-steps/s" and info["kind"] == "port" and info["value"] > 0 and info["cores"] >= 1
    assert "cartpole_vpg_65536x200" in info["sample"]

"""bench.py contract on the CPU: the reference arm (CPU oracle on the host cores) prints ONE JSON line with the keys the
driver reads; the GPU arm refuses to run without a device instead of falling back."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, cwd=ROOT, timeout=600)


def test_reference_arm_json_line():
    for wl, shape in (("c1", "752x480"), ("c4", "640x480")):
        p = run("--impl", "reference", "--steps", "1", "--warmup", "0", "--workload", wl)
        assert p.returncode == 0, p.stderr[-500:]
        lines = [l for l in p.stdout.splitlines() if l.strip()]
        assert len(lines) == 1
        d = json.loads(lines[0])
        assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True and d["value"] > 0
        assert shape in d["metric"] and shape in d["config"]["workload"]
        assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
        assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
        assert d["gpu_launches"] == 0 and d["vs_baseline"] is None and d["dtype"] == "u8"


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, cwd=ROOT, env=env, timeout=120)
    assert p.returncode == 0 and p.stdout.strip() == ""

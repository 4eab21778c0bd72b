"""CPU-only: the reference arm of bench.py prints the contract's JSON line (the GPU arm needs a B200)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(*args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    return json.loads(out.stdout.strip().splitlines()[-1])


def test_reference_arm_line_cfg1(oracle, sla):
    d = run("--impl", "reference", "--workload", "cfg1", "--steps", "2", "--warmup", "1")
    assert d["impl"] == "reference" and d["metric"] == "bid_arcs_per_sec" and d["unit"] == "bid-arcs/s"
    assert d["higher_is_better"] is True and d["steps"] == 2 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert "cfg1" in d["config"]["workload"] and "model" not in d["config"]


def test_reference_arm_other_ranks_stay_silent(oracle, sla):
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cfg1", "--gpus", "2",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=120, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""

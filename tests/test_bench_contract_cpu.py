"""bench.py --impl reference on a CPU-only box: the reference arm of the bench contract (oracle port of the reference's CPU
operator apply, global_curved.jl:470-492) prints one JSON line with the contract's keys; under torchrun every rank but 0
exits without work."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    # 64 x 64-point blocks keep the oracle's assembly short; the contract is the same at the default size
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
           "--n", "63", "--cpu-blocks", "2"]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)


def test_reference_arm_prints_the_contract_line():
    r = run()
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "GDOF/s" and d["higher_is_better"] is True
    assert d["metric"] == "fp64 SBP operator-apply GDOF/s" and d["dtype"] == "f64" and d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": "GDOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["gpu_launches"] == 0 and d["steps"] == 2 and d["warmup"] == 1
    assert "workload" in d["config"]


def test_reference_arm_other_ranks_exit_without_work():
    r = run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == "", (r.stdout, r.stderr[-500:])

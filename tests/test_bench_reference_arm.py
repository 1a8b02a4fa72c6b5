"""bench.py --impl reference: the reference's CPU path through the bench contract (no GPU involved)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra=None):
    env = dict(os.environ, **(env_extra or {}))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                        "--chains", "64", "--points", "1000"], capture_output=True, text=True, env=env, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    return [l for l in r.stdout.splitlines() if l.startswith("{")]


def test_reference_arm_prints_one_contract_line():
    lines = _run()
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "model_point_logL_evals_per_sec" and line["unit"] == "points/s"
    assert line["value"] > 0 and line["higher_is_better"] is True and line["steps"] == 2 and line["dtype"] == "f64"
    assert line["gpu_launches"] == 0 and line["vs_baseline"] is None
    assert line["e2e"] == {"value": line["value"], "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert "workload" in line["config"] and "model" not in line["config"]


def test_reference_arm_other_ranks_do_nothing():
    # under torchrun rank 0 alone runs the CPU reference; the other ranks exit 0 without work or output
    assert _run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}) == []

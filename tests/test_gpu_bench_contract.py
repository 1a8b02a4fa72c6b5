"""bench.py on the B200: one JSON line carrying the contract's keys (small sizes; the numbers are not judged here)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bench_line_has_the_contract_keys():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "3", "--warmup", "3", "--chains", "256",
                        "--points", "4000", "--pt-steps", "4", "--no-pt-reference"], capture_output=True, text=True,
                       timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["metric"] == "model_point_logL_evals_per_sec" and line["unit"] == "points/s" and line["n_gpus"] == 1
    assert line["value"] > 0 and line["steps"] == 3 and line["warmup"] >= 3 and line["scaling"] == "weak"
    assert line["dtype"] == "f64" and line["data"] == "synthetic" and line["gpu_launches"] >= 2 * 3
    e2e = line["e2e"]
    assert e2e["value"] > 0 and e2e["h2d_bytes_per_step"] == 256 * 21 * 8
    assert e2e["d2h_bytes_per_step"] == 256 * 8
    roof = line["roofline"]
    assert roof["kernel"] == "k_chain_eval" and roof["peak"] > 0 and roof["achieved"] > 0
    assert abs(roof["frac"] - roof["achieved"] / roof["peak"]) < 1e-12 and 0 < roof["kernel_share_of_step"] <= 1.0
    assert roof["traffic"] is None and roof["executed"] is None  # both belong to the C2 capture only
    clocks = line["clocks"]
    assert clocks["sm_mhz"] > 0 and clocks["sm_max_mhz"] >= clocks["sm_mhz"] and isinstance(clocks["reasons"], list)
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["value"] > 0 and cb["cores"] >= 1
    assert cb["max_rel_err_gpu_vs_cpu_on_sample"] <= 1e-10
    assert roof["frac"] <= 1.0 and roof["frac_vs_reference_formulation"] > roof["frac"] and roof["flop_per_point"] < 520
    pt = line["pt"]
    assert pt["steps_per_sec"] > 0 and pt["cold_logL_finite"] is True
    # walkers are counted by the likelihood kernel itself: never more than were proposed
    assert 0 < pt["evaluated_walkers_per_step"] <= pt["walkers"] and 0 <= pt["skipped_fraction"] < 1
    assert pt["model_points_per_sec"] <= pt["walkers"] * pt["n_points"] * pt["steps_per_sec"] * (1 + 1e-9)
    extra = line["extra"]
    for key in ("C1_latency", "C3_share", "C4", "C5_share"):
        assert extra[key]["ms"] > 0 and extra[key]["nan_fraction"] < 1e-3  # (e -> 0.95 draws may be NaN in the reference too)
    assert extra["C4"]["n_chains"] == 8192 and extra["C4"]["n_points"] == 50000
    assert extra["C5_share"]["n_chains"] == 2048 and extra["C5_share"]["n_points"] == 200000
    assert extra["pt_one_ladder_64_rungs_x_200k"]["steps_per_sec"] > 0

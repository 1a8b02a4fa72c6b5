"""Writes tests/golden/outliers_v1.json: the chains of the seeded parity campaign (12 passes, SEED_OFFSET 5000,
50 688 chains; tests/parity_scan_lib.py) whose GPU logL differs from the compiled reference's by more than the
1e-10 gate, with the parameters, the data-set key and the reference value as hex floats.  Needs a B200:

    python tests/golden/make_outliers.py [out.json]

The fixtures are read by tests/test_gpu_parity_scan.py (GPU, amended bound) and tests/test_outlier_fixtures.py
(CPU: the reference against itself compiled with FMA contraction on the same chains)."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402
import hb_mcmc_b200 as hb  # noqa: E402
import parity_scan_lib as ps  # noqa: E402

R = oracle.Reference() if oracle.have_reference() else oracle.Oracle()
ctx = hb.Context(0)
rel, rec = ps.scan(ctx, R, range(12), seed_offset=5000, report=lambda s: print(s, flush=True))
out = [r for r in rec if r["rel"] > 1e-10]
print("WORST", rel.max(), "chains", rel.size, "above 1e-10:", len(out), "above 5e-11:", len(rec))
path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "outliers_v1.json")
json.dump(out, open(path, "w"), indent=1)
json.dump(rec, open(path.replace(".json", "_above_5e-11.json"), "w"), indent=1)

# wide parity scan vs the compiled reference: many chains, several N / truths / e ranges
#   REPS=12 SEED_OFFSET=5000 python tests/tools/parity_scan.py [outliers.json]
import json, os, sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, oracle
import hb_mcmc_b200 as hb
import parity_scan_lib as ps
R = oracle.Reference() if oracle.have_reference() else oracle.Oracle()
if os.environ.get("HB_LIB"):
    from hb_mcmc_b200 import lib as hblib
    hblib._lib = hblib.load_library(os.environ["HB_LIB"])
ctx = hb.Context(0)
reps = range(int(os.environ.get("REPS", "1")))
rel, rec = ps.scan(ctx, R, reps, int(os.environ.get("SEED_OFFSET", "0")), report=lambda s: print(s, flush=True))
print("WORST", rel.max(), "chains", rel.size, "above 1e-10:", int((rel > 1e-10).sum()), "above 5e-11:", len(rec))
if len(sys.argv) > 1:
    json.dump(rec, open(sys.argv[1], "w"), indent=1)

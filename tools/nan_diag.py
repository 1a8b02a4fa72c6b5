"""Where does the NaN pattern of the CUDA likelihood differ from the compiled reference?  (diagnosis tool:
the seeded scan of tests/test_gpu_parity_scan.py, chain by chain, the in-tree library next to HB_LIB_B)"""
import os, sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, oracle
import hb_mcmc_b200 as hb
from hb_mcmc_b200 import workload as wl
import parity_scan_lib as ps
R = oracle.Reference() if oracle.have_reference() else oracle.Oracle()
ctx = hb.Context(0)
for rep in range(8, 13):
    for truth_name, N, emax, n, seed0 in ps.SETS:
        seed = seed0 + 100 * rep + 5000
        truth = ps.TRUTHS[truth_name]
        t, fl, er = wl.make_dataset(N, truth, R.calc_light_curve)
        ctx.set_data(t, fl, er)
        P = ps.chain_set(truth_name, N, emax, n, seed, ctx.roche_overflow)
        g = ctx.loglikelihood(P)
        o = R.loglikelihood_batch(t, fl, er, P)
        bad = np.nonzero(np.isnan(g) != np.isnan(o))[0]
        with np.errstate(invalid="ignore", divide="ignore"):
            rel = np.where(np.isnan(o) | np.isnan(g), 0.0, np.abs(g - o) / np.abs(o))
        print(f"set {truth_name} N={N} seed={seed}: n={len(P)} nan-mismatch {len(bad)} max rel {rel.max():.2e}", flush=True)
        for i in bad:
            print("   chain", i, "e =", P[i, 3], "gpu", g[i], "ref", o[i], flush=True)
            lc_g = ctx.light_curves(P[i:i + 1])[0]
            lc_r = R.calc_light_curve(t, P[i])
            print("   gpu lc nan count", int(np.isnan(lc_g).sum()), " ref lc nan count", int(np.isnan(lc_r).sum()),
                  " ref nan idx", np.nonzero(np.isnan(lc_r))[0][:5], flush=True)
            print("   params", [float(v).hex() for v in P[i]], flush=True)

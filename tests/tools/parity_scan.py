# wide parity scan vs the compiled reference: many chains, several N / truths / e ranges
import sys; sys.path.insert(0, ".")
import numpy as np, oracle
import hb_mcmc_b200 as hb
from hb_mcmc_b200 import workload as wl
import os
R = oracle.Reference()
if os.environ.get("HB_LIB"):
    from hb_mcmc_b200 import lib as hblib
    hblib._lib = hblib.load_library(os.environ["HB_LIB"])
ctx = hb.Context(0)
worst = 0
REPS = int(os.environ.get("REPS", "1"))
over = 0
total = 0
for rep in range(REPS):
  for truth, N, emax, n, seed0 in ((wl.TRUTH_A, 20000, 0.95, 512, 1), (wl.TRUTH_B, 20000, 0.99, 512, 2), (wl.TRUTH_A, 1001, 0.99, 1024, 3), (wl.TRUTH_B, 50000, 0.97, 128, 4), (wl.TRUTH_A, 375, 0.9, 2048, 5)):
    seed = seed0 + 100 * rep + int(os.environ.get("SEED_OFFSET", "0"))
    t, fl, er = wl.make_dataset(N, truth, R.calc_light_curve)
    ctx.set_data(t, fl, er)
    P = wl.draw_chains(n, truth, ctx.roche_overflow, seed=seed, e_max=emax)
    P[0] = truth
    # half the chains near the truth (chi^2 ~ N: the most sensitive regime)
    k = n // 2
    P[1:k] = truth + 1e-3 * np.random.default_rng(seed).standard_normal((k - 1, 21)) * np.abs(truth + 0.1)
    P[1:k, 2] = truth[2]
    keep = ctx.roche_overflow(P) == 0
    P = P[keep]
    g = ctx.loglikelihood(P); o = R.loglikelihood_batch(t, fl, er, P)
    rel = np.abs(g - o) / np.abs(o)
    assert np.array_equal(np.isnan(g), np.isnan(o))
    i = np.nanargmax(rel)
    print(f"truth e={truth[3]:.3f} N={N:6d} n={len(P):5d} emax={emax}: max rel {np.nanmax(rel):.3e} (e={P[i,3]:.3f}, logL={o[i]:.4g})  median {np.nanmedian(rel):.2e}  nan {np.isnan(o).sum()}")
    worst = max(worst, np.nanmax(rel))
    over += int(np.nansum(rel > 1e-10))
    total += len(P)
print("WORST", worst, "chains", total, "above 1e-10:", over)

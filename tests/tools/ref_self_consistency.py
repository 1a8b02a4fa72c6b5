# How reproducible is the reference's own logL?  The chain sets of parity_scan.py, evaluated by the compiled
# reference (README flags: ISO C, no contraction) and by the SAME unmodified file compiled with FMA contraction
# (`make -C oracle ref_fma`).  CPU only.  The size and the location of the differences (in-eclipse samples next to
# d = sqrt(R1^2 - R2^2), likelihood3.c:372-376) are what the GPU-vs-reference outliers of DESIGN.md section 4 are.
import os, sys; sys.path.insert(0, ".")
import numpy as np, oracle
from hb_mcmc_b200 import workload as wl
R, F = oracle.Reference(), oracle.Reference(variant="fma")
roche = lambda P: np.array([R.roche_overflow(p) for p in np.atleast_2d(P)])
REPS = int(os.environ.get("REPS", "1"))
worst, over, total, nan_mismatch = 0.0, 0, 0, 0
for rep in range(REPS):
  for truth, N, emax, n, seed0 in ((wl.TRUTH_A, 20000, 0.95, 512, 1), (wl.TRUTH_B, 20000, 0.99, 512, 2), (wl.TRUTH_A, 1001, 0.99, 1024, 3), (wl.TRUTH_B, 50000, 0.97, 128, 4), (wl.TRUTH_A, 375, 0.9, 2048, 5)):
    seed = seed0 + 100 * rep + int(os.environ.get("SEED_OFFSET", "0"))
    t, fl, er = wl.make_dataset(N, truth, R.calc_light_curve)
    P = wl.draw_chains(n, truth, roche, seed=seed, e_max=emax)
    P[0] = truth
    k = n // 2
    P[1:k] = truth + 1e-3 * np.random.default_rng(seed).standard_normal((k - 1, 21)) * np.abs(truth + 0.1)
    P[1:k, 2] = truth[2]
    P = P[roche(P) == 0]
    a, b = R.loglikelihood_batch(t, fl, er, P), F.loglikelihood_batch(t, fl, er, P)
    mism = int(np.sum(np.isnan(a) != np.isnan(b)))  # e -> 1: whether the un-converged solve ends in NaN can flip too
    nan_mismatch += mism
    rel = np.abs(a - b) / np.abs(a)
    print(f"truth e={truth[3]:.3f} N={N:6d} n={len(P):5d}: reference vs reference+FMA max rel {np.nanmax(rel):.3e}  median {np.nanmedian(rel):.2e}  NaN on one side only: {mism}", flush=True)
    worst = max(worst, np.nanmax(rel)); over += int(np.nansum(rel > 1e-10)); total += len(P)
print("WORST", worst, "chains", total, "above 1e-10:", over, "NaN on one side only:", nan_mismatch)

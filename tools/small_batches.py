"""Batches smaller than the grid, three calls each (for the ncu launch list of tools/prof_r2.sh): C1, a 50-rung ladder of
20 000-point light curves, 8 chains of 200 000 points."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import hb_mcmc_b200 as hb  # noqa: E402
from hb_mcmc_b200 import workload as wl  # noqa: E402

ctx = hb.Context(0)
for n, N, truth in ((1, 20000, wl.TRUTH_A), (50, 20000, wl.TRUTH_A), (8, 200000, wl.TRUTH_B)):
    t, flux, err = wl.make_dataset(N, truth, ctx.calc_light_curve)
    ctx.set_data(t, flux, err)
    P = wl.draw_chains(n, truth, ctx.roche_overflow, seed=1)
    ctx.loglikelihood(P)
    t0 = time.perf_counter()
    for _ in range(3):
        ctx.loglikelihood(P)
    print(f"{n} x {N}: {(time.perf_counter() - t0) / 3 * 1e6:.1f} us per host-buffer call", flush=True)

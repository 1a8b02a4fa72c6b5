"""Reference-size sampler (one ladder of 50 rungs on the real folded light curve TIC 102289966, 375 points): steps/s of
the one-launch step loop (k_pt_run) against the stream-ordered kernels replayed from CUDA graphs.  Needs a B200."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import hb_mcmc_b200 as hb  # noqa: E402
from hb_mcmc_b200.pt import PTSampler  # noqa: E402

ctx = hb.Context(0)
d = np.loadtxt("tests/golden/lc_102289966_new.txt", skiprows=1)
ctx.set_data(d[:, 0], d[:, 1], d[:, 2])
for n_temps, n_ens in ((50, 1), (50, 8), (64, 9)):
    for one in (True, False):
        s = PTSampler(ctx, n_temps, n_ens, 0.7960497, seed=3)
        s.set_one_launch(one)
        s.init_random()
        s.step(600)
        ctx.sync()
        t0 = time.perf_counter()
        s.step(3000)
        ctx.sync()
        dt = time.perf_counter() - t0
        print(f"{n_temps} rungs x {n_ens} ladders x {len(d)} points, {'one launch' if one else 'stream-ordered, graphs'}: "
              f"{3000 / dt:9.0f} steps/s ({dt / 3000 * 1e6:.1f} us per step)", flush=True)
        s.close()

import sys, time
import numpy as np
sys.path.insert(0, ".")
import hb_mcmc_b200 as hb
from hb_mcmc_b200 import workload as wl
from hb_mcmc_b200.pt import PTSampler
ctx = hb.Context(0)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
T, E = int(sys.argv[2]) if len(sys.argv) > 2 else 64, int(sys.argv[3]) if len(sys.argv) > 3 else 32
nstep = int(sys.argv[4]) if len(sys.argv) > 4 else 5
t, flux, err = wl.make_dataset(N, wl.TRUTH_A, ctx.calc_light_curve)
ctx.set_data(t, flux, err)
s = PTSampler(ctx, T, E, float(wl.TRUTH_A[2]), seed=11)
s.init_random()
s.step(3); ctx.sync()
t0 = time.time(); s.step(nstep); ctx.sync(); dt = time.time() - t0
print(f"N={N} T={T} E={E}: {dt/nstep*1e3:.3f} ms/step  {nstep/dt:.1f} steps/s")
x, ll, idx = s.state()
print("nan frac", np.isnan(ll).mean(), "roche frac", (ll == -5e14).mean(), "e>0.9 frac", (x[:,3] > 0.9).mean())

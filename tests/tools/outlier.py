import sys; sys.path.insert(0, ".")
import numpy as np, oracle
import hb_mcmc_b200 as hb
from hb_mcmc_b200 import workload as wl
R = oracle.Reference(); O = oracle.Oracle(); ctx = hb.Context(0)
truth, N, emax, n, seed = wl.TRUTH_A, 20000, 0.95, 512, 1
t, fl, er = wl.make_dataset(N, truth, R.calc_light_curve)
ctx.set_data(t, fl, er)
P = wl.draw_chains(n, truth, ctx.roche_overflow, seed=seed, e_max=emax)
P[0] = truth
k = n // 2
P[1:k] = truth + 1e-3 * np.random.default_rng(seed).standard_normal((k - 1, 21)) * np.abs(truth + 0.1)
P[1:k, 2] = truth[2]
P = P[ctx.roche_overflow(P) == 0]
g = ctx.loglikelihood(P); o = R.loglikelihood_batch(t, fl, er, P)
rel = np.abs(g - o) / np.abs(o)
for i in np.argsort(-rel)[:4]:
    lc_g = ctx.calc_light_curve(t, P[i]); lc_o, raw_o = O.calc_light_curve(t, P[i], raw=True)
    d = lc_g - lc_o
    j = np.argmax(np.abs(d))
    srt = np.sort(raw_o); kr = 10000
    print(f"chain {i}: rel {rel[i]:.2e} logL {o[i]:.6f} dlogL {g[i]-o[i]:.3e}; lc max|d| {np.abs(d).max():.2e} at {j} (lc={lc_o[j]:.4f}); mean d {d.mean():.3e} median d {np.median(d):.3e}; n(|d|>1e-14) {(np.abs(d)>1e-14).sum()}; gap at median {srt[kr+1]-srt[kr]:.2e} {srt[kr]-srt[kr-1]:.2e}")
    # chi2 from the two light curves in numpy
    w = 1/er
    print("   chi2(lc_g) - chi2(lc_o) =", np.sum(((lc_g-fl)*w)**2) - np.sum(((lc_o-fl)*w)**2), " -2*dlogL =", -2*(g[i]-o[i]))

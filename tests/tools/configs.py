# all BASELINE configs on one GPU (per-GPU share for the multi-GPU ones): time + parity on a subset
import sys, time
import numpy as np
sys.path.insert(0, ".")
import oracle, torch
import hb_mcmc_b200 as hb
from hb_mcmc_b200 import workload as wl
O = oracle.Oracle()
try: R = oracle.Reference()
except Exception: R = O
ctx = hb.Context(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
for name, n, N, truth, gaia in (("C1", 1, 20000, wl.TRUTH_A, False), ("C2", 4096, 20000, wl.TRUTH_A, False), ("C3/8", 2048, 20000, wl.TRUTH_A, False),
                                ("C4", 8192, 50000, wl.TRUTH_A, True), ("C5/8", 2048, 200000, wl.TRUTH_B, False), ("real", 4096, 375, wl.TRUTH_A, False)):
    t, flux, err = wl.make_dataset(N, truth, ctx.calc_light_curve)
    ctx.set_data(t, flux, err)
    md, me = [1000, 1, 1, 1, 1], [1e15] * 4
    if gaia:
        G = ctx.chain_info(truth[None], 100.0)[0, 4]
        md, me = [100.0, G + 0.02, 1, 1, 1], [0.05, 1e15, 1e15, 1e15]
    ctx.set_mags(md, me, 1, 0)
    P = wl.draw_chains(n, truth, ctx.roche_overflow, seed=1)
    P[0] = truth
    dP = torch.from_numpy(P).cuda(); dL = torch.empty(n, dtype=torch.float64, device="cuda")
    for _ in range(2): ctx.loglikelihood_dev(dP.data_ptr(), n, dL.data_ptr())
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    reps = 5 if n * N > 1e8 else 20
    e0.record(stream)
    for _ in range(reps): ctx.loglikelihood_dev(dP.data_ptr(), n, dL.data_ptr())
    e1.record(stream); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    got = dL.cpu().numpy()
    k = min(n, 12 if N > 50000 else 32)
    want = R.loglikelihood_batch(t, flux, err, P[:k], md, me)
    rel = np.abs(got[:k] - want) / np.abs(want)
    print(f"{name:5s} {n:5d} x {N:6d}: {ms:9.3f} ms  {n*N/ms*1e3:.3e} pts/s  nan {np.isnan(got).mean():.4f}  parity max rel {np.nanmax(rel):.2e} (n={k}) logL[0]={got[0]:.6f}", flush=True)

import sys, glob, time, os
import numpy as np
sys.path.insert(0, ".")
import hb_mcmc_b200 as hb
from hb_mcmc_b200 import lib as hblib, workload as wl
import torch

def make_ctx(path):
    L = hblib.load_library(path)
    c = hb.Context.__new__(hb.Context)
    import ctypes as C
    h = C.c_void_p()
    rc = L.hb_create(C.byref(h), 0); assert rc == 0
    c._L = L; c._h = h; c.device = 0; c.n_points = 0
    return c

N = int(os.environ.get("NPTS", 20000)); n = int(os.environ.get("NCHAINS", 4096))
base = hb.Context(0)
t, flux, err = wl.make_dataset(N, wl.TRUTH_A, base.calc_light_curve)
P = wl.draw_chains(n, wl.TRUTH_A, base.roche_overflow, seed=1)
base.set_data(t, flux, err)
ref = base.loglikelihood(P)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
dP = torch.from_numpy(P).cuda(); dL = torch.empty(n, dtype=torch.float64, device="cuda")
for path in [None] + sorted(glob.glob("tools/variants/lib_*.so")):
    c = base if path is None else make_ctx(path)
    if path is not None: c.set_data(t, flux, err)
    c.set_stream(stream.cuda_stream)
    for _ in range(3): c.loglikelihood_dev(dP.data_ptr(), n, dL.data_ptr())
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record(stream)
    for _ in range(reps): c.loglikelihood_dev(dP.data_ptr(), n, dL.data_ptr())
    e1.record(stream); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    got = dL.cpu().numpy()
    rel = np.nanmax(np.abs(got - ref) / np.abs(ref))
    print(f"{(path or 'default'):45s} {ms:7.3f} ms  {n*N/ms*1e3:.3e} pts/s   max rel vs default {rel:.2e}", flush=True)
    c.set_stream(None)

# time the likelihood for batches of chains at a fixed eccentricity (default library and tools/variants/*)
import sys, glob, os
import numpy as np
sys.path.insert(0, ".")
import hb_mcmc_b200 as hb
from hb_mcmc_b200 import lib as hblib, workload as wl
import torch, ctypes as C

def make_ctx(path):
    L = hblib.load_library(path)
    c = hb.Context.__new__(hb.Context)
    h = C.c_void_p()
    assert L.hb_create(C.byref(h), 0) == 0
    c._L = L; c._h = h; c.device = 0; c.n_points = 0
    return c

N, n = 20000, 4096
base = hb.Context(0)
t, flux, err = wl.make_dataset(N, wl.TRUTH_A, base.calc_light_curve)
P0 = wl.draw_chains(n, wl.TRUTH_A, base.roche_overflow, seed=1)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
for path in [None] + sorted(glob.glob("tools/variants/lib_*.so")):
    c = base if path is None else make_ctx(path)
    c.set_data(t, flux, err); c.set_stream(stream.cuda_stream)
    for e in (0.3, 0.7, 0.82, 0.9, 0.95):
        P = P0.copy(); P[:, 3] = e
        P = P[base.roche_overflow(P) == 0]
        dP = torch.from_numpy(P).cuda(); dL = torch.empty(len(P), dtype=torch.float64, device="cuda")
        for _ in range(3): c.loglikelihood_dev(dP.data_ptr(), len(P), dL.data_ptr())
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(10): c.loglikelihood_dev(dP.data_ptr(), len(P), dL.data_ptr())
        e1.record(stream); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"{(path or 'default'):32s} e={e:.2f} chains={len(P):5d} {ms:7.3f} ms  {len(P)*N/ms*1e3:.3e} pts/s", flush=True)
    c.set_stream(None)

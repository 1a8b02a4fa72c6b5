"""A/B timing of libhb_b200 builds on one GPU: the in-tree library against every tools/variants/lib_*.so (older
commits, -D variants), same inputs, device-resident, CUDA events; also checks that results agree.
    NPTS=20000 NCHAINS=4096 python tools/ab.py            (needs a B200)"""
import ctypes as C
import glob
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from hb_mcmc_b200 import workload as wl  # noqa: E402

dp = C.POINTER(C.c_double)


class Lib:
    def __init__(self, path):
        L = self.L = C.CDLL(path)
        vp, l = C.c_void_p, C.c_long
        L.hb_create.argtypes = [C.POINTER(vp), C.c_int]
        L.hb_set_data.argtypes = [vp, dp, dp, dp, l]
        L.hb_set_stream.argtypes = [vp, vp]
        L.hb_loglikelihood_batch_dev.argtypes = [vp, vp, l, vp]
        L.hb_calc_light_curve.argtypes = [vp, dp, l, dp, dp]
        L.hb_chain_info_batch.argtypes = [vp, dp, l, C.c_double, dp]
        L.hb_last_error.argtypes = [vp]
        L.hb_last_error.restype = C.c_char_p
        self.h = vp()
        assert L.hb_create(C.byref(self.h), 0) == 0

    def ck(self, rc):
        assert rc == 0, self.L.hb_last_error(self.h)

    def calc_light_curve(self, t, p):
        t, p = np.ascontiguousarray(t), np.ascontiguousarray(p)
        out = np.empty(t.size)
        self.ck(self.L.hb_calc_light_curve(self.h, t.ctypes.data_as(dp), t.size, p.ctypes.data_as(dp), out.ctypes.data_as(dp)))
        return out

    def roche_overflow(self, P):
        P = np.ascontiguousarray(P).reshape(-1, 21)
        out = np.empty((P.shape[0], 9))
        self.ck(self.L.hb_chain_info_batch(self.h, P.ctypes.data_as(dp), P.shape[0], 1000.0, out.ctypes.data_as(dp)))
        return out[:, 8].astype(np.int32)

    def set_data(self, t, f, e):
        t, f, e = (np.ascontiguousarray(x) for x in (t, f, e))
        self.ck(self.L.hb_set_data(self.h, t.ctypes.data_as(dp), f.ctypes.data_as(dp), e.ctypes.data_as(dp), t.size))


N = int(os.environ.get("NPTS", 20000))
n = int(os.environ.get("NCHAINS", 4096))
truth = wl.TRUTH_B if os.environ.get("TRUTH", "A") == "B" else wl.TRUTH_A
paths = ["hb_mcmc_b200/csrc/libhb_b200.so"] + sorted(glob.glob("tools/variants/lib_*.so"))
base = Lib(paths[0])
t, flux, err = wl.make_dataset(N, truth, base.calc_light_curve)
P = wl.draw_chains(n, truth, base.roche_overflow, seed=1)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
dP = torch.from_numpy(P).cuda()
dL = torch.empty(n, dtype=torch.float64, device="cuda")
ref = None
for rnd in range(int(os.environ.get("ROUNDS", 2))):
    for path in paths:
        c = base if path == paths[0] else Lib(path)
        c.set_data(t, flux, err)
        c.ck(c.L.hb_set_stream(c.h, C.c_void_p(stream.cuda_stream)))
        for _ in range(3):
            c.ck(c.L.hb_loglikelihood_batch_dev(c.h, C.c_void_p(dP.data_ptr()), n, C.c_void_p(dL.data_ptr())))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = max(3, min(50, int(2e9 / (n * N))))
        e0.record(stream)
        for _ in range(reps):
            c.ck(c.L.hb_loglikelihood_batch_dev(c.h, C.c_void_p(dP.data_ptr()), n, C.c_void_p(dL.data_ptr())))
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        got = dL.cpu().numpy()
        if ref is None:
            ref = got.copy()
        rel = np.nanmax(np.abs(got - ref) / np.abs(ref))
        print(f"{os.path.basename(path):40s} {n:6d} x {N:6d}: {ms*1e3:10.1f} us  {n*N/ms*1e3:.3e} pts/s   max rel vs first {rel:.2e}", flush=True)
        c.ck(c.L.hb_set_stream(c.h, None))

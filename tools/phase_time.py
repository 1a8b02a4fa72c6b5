"""Per-chain phase cycles of k_chain_eval from a -DHB_PHASE_PROF build (tools/build_variants.sh prof -DHB_PHASE_PROF): table, pre-sample, model pass,
hand-over, select.   python tools/phase_time.py tools/variants/lib_prof*.so"""
import ctypes as C, sys
import numpy as np, torch
sys.path.insert(0, ".")
from hb_mcmc_b200 import workload as wl
sys.argv = [sys.argv[0]] + sys.argv[1:]
import importlib.util
spec = importlib.util.spec_from_file_location("ab", "tools/ab.py")
dp = C.POINTER(C.c_double)
N, n = 20000, 4096
for path in sys.argv[1:]:
    L = C.CDLL(path)
    vp, l = C.c_void_p, C.c_long
    L.hb_create.argtypes = [C.POINTER(vp), C.c_int]
    L.hb_set_data.argtypes = [vp, dp, dp, dp, l]
    L.hb_loglikelihood_batch_dev.argtypes = [vp, vp, l, vp]
    L.hb_calc_light_curve.argtypes = [vp, dp, l, dp, dp]
    L.hb_chain_info_batch.argtypes = [vp, dp, l, C.c_double, dp]
    L.hb_phase_read.argtypes = [C.POINTER(C.c_ulonglong)]
    h = vp(); assert L.hb_create(C.byref(h), 0) == 0
    def lc(t, p):
        t, p = np.ascontiguousarray(t), np.ascontiguousarray(p); out = np.empty(t.size)
        assert L.hb_calc_light_curve(h, t.ctypes.data_as(dp), t.size, p.ctypes.data_as(dp), out.ctypes.data_as(dp)) == 0
        return out
    def roche(P):
        P = np.ascontiguousarray(P).reshape(-1, 21); out = np.empty((P.shape[0], 9))
        assert L.hb_chain_info_batch(h, P.ctypes.data_as(dp), P.shape[0], 1000.0, out.ctypes.data_as(dp)) == 0
        return out[:, 8].astype(np.int32)
    t, fl, er = wl.make_dataset(N, wl.TRUTH_A, lc)
    P = wl.draw_chains(n, wl.TRUTH_A, roche, seed=1)
    t, fl, er = (np.ascontiguousarray(x) for x in (t, fl, er))
    assert L.hb_set_data(h, t.ctypes.data_as(dp), fl.ctypes.data_as(dp), er.ctypes.data_as(dp), t.size) == 0
    dP = torch.from_numpy(P).cuda(); dL = torch.empty(n, dtype=torch.float64, device="cuda")
    ph = (C.c_ulonglong * 8)()
    for rep in range(3):
        assert L.hb_loglikelihood_batch_dev(h, C.c_void_p(dP.data_ptr()), n, C.c_void_p(dL.data_ptr())) == 0
        torch.cuda.synchronize()
        L.hb_phase_read(ph)
    v = np.array(list(ph)[:5], dtype=np.float64) / n
    print(f"{path}: cycles per chain  table {v[0]:.0f}  pre-sample {v[1]:.0f}  pass {v[2]:.0f}  post {v[3]:.0f}  select {v[4]:.0f}", flush=True)

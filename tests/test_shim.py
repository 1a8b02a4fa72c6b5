"""libhb_likelihood3.so: the reference's likelihood3.h symbols on top of the CUDA library."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from hb_mcmc_b200 import build
from hb_mcmc_b200 import workload as wl

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
dp = C.POINTER(C.c_double)
REF_SRC = "/root/reference/src"


@pytest.fixture(scope="module")
def shim():
    build.build_lib()
    path = build.build_shim()
    L = C.CDLL(path)
    L.loglikelihood.restype = C.c_double
    L.loglikelihood.argtypes = [dp, dp, dp, C.c_long, dp, dp, dp]
    L.calc_light_curve.argtypes = [dp, C.c_long, dp, dp]
    L.calc_mags.argtypes = [dp, C.c_double, dp, dp, dp, dp]
    L.calc_radii_and_Teffs.argtypes = [dp] * 5
    L.RocheOverflow.restype = C.c_int
    L.RocheOverflow.argtypes = [dp]
    L.remove_median.argtypes = [dp, C.c_long, C.c_long]
    L.traj.argtypes = [dp] * 7 + [C.c_int]
    for n in ("_getT", "_getR", "envelope_Temp", "envelope_Radius", "get_alpha_beam"):
        getattr(L, n).restype = C.c_double
        getattr(L, n).argtypes = [C.c_double]
    L.eclipse_area.restype = C.c_double
    L.eclipse_area.argtypes = [C.c_double] * 3
    L.set_limits.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double]
    L.initialize_proposals.argtypes = [dp, C.c_void_p]
    L.quickSort.argtypes = [dp, C.c_int, C.c_int]
    L.hb_shim_set_memo.argtypes = [C.c_int]
    L.hb_shim_set_flags.argtypes = [C.c_int, C.c_int]
    L.hb_shim_memo_hits.restype = C.c_long
    return L


def p(a):
    return a.ctypes.data_as(dp)


def test_exports_every_declared_symbol(shim):
    src = open(os.path.join(ROOT, "include", "hb_likelihood3_abi.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = set(re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", src)) - {"defined"}
    assert len(names) >= 22
    for n in names:
        assert hasattr(shim, n), n


def test_host_tables_match_golden(shim, golden):
    limited = (C.c_double * 42)()
    limits = (C.c_double * 42)()
    gauss = (C.c_int * 21)()
    shim.set_limits(limited, limits, gauss, 2.0)
    ld, lm = np.array(limited).reshape(21, 2), np.array(limits).reshape(21, 2)
    got = np.stack([lm[:, 0], lm[:, 1], ld[:, 0], ld[:, 1], np.array(gauss, dtype=np.float64)])
    assert np.array_equal(got, golden["limits_P2"])
    sig = np.zeros(21)
    shim.initialize_proposals(p(sig), None)
    assert np.array_equal(sig, golden["sigmas"])
    x = np.random.default_rng(0).standard_normal(1000)
    y = x.copy()
    shim.quickSort(p(y), 0, 999)
    assert np.array_equal(y, np.sort(x))


def test_reference_driver_links_against_shim(tmp_path):
    """The UNMODIFIED mcmc_wrapper2.c compiles and links with likelihood3.c replaced by the shim."""
    if not os.path.exists(os.path.join(REF_SRC, "mcmc_wrapper2.c")):
        pytest.skip("reference sources not present")
    build.build_lib()
    build.build_shim()
    out = tmp_path / "hb_mcmc_ref_driver"
    cmd = ["gcc", "-O3", "-std=c99", "-fopenmp", "-w", os.path.join(REF_SRC, "mcmc_wrapper2.c"), "-o", str(out),
           "-L", build.CSRC, "-lhb_likelihood3", "-lhb_b200", f"-Wl,-rpath,{build.CSRC}", "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    nm = subprocess.run(["nm", "-D", "--undefined-only", str(out)], capture_output=True, text=True).stdout
    for sym in ("loglikelihood", "calc_light_curve", "set_limits", "initialize_proposals"):
        assert re.search(rf"\bU {sym}\b", nm), sym


@pytest.mark.gpu
def test_shim_matches_golden(shim, golden):
    kat = golden["kat_params"].copy()
    t = golden["kat_times"].copy()
    lc = np.empty(1000)
    shim.calc_light_curve(p(t), 1000, p(kat), p(lc))
    assert np.abs(lc - golden["kat_lc"]).max() < 2e-12
    R = [C.c_double() for _ in range(4)]
    shim.calc_radii_and_Teffs(p(kat), *[C.byref(r) for r in R])
    assert np.allclose([r.value for r in R], golden["kat_radii_teffs"], rtol=1e-13, atol=0)
    m = [C.c_double() for _ in range(4)]
    shim.calc_mags(p(kat), 100.0, *[C.byref(r) for r in m])
    assert np.abs(np.array([r.value for r in m]) - golden["kat_mags_D100"]).max() < 1e-12
    assert shim.RocheOverflow(p(kat)) == 0
    assert abs(shim.get_alpha_beam(3.8) - 0.81250000000000044) < 1e-14
    assert abs(shim.eclipse_area(1.0, 0.5, 0.8 * 6.955e10) - 0.54910621859670772) < 1e-13
    assert abs(shim._getT(0.3) - golden["getT"][np.argmin(np.abs(golden["logM_grid"] - 0.3))]) < 1e-12
    # loglikelihood: same arrays, in-place clamp of noise[] (quirk Q2), Gaia term from the arguments
    flux, err = np.ones(1000), np.full(1000, 1e-3)
    md, me = golden["kat_mag_data"].copy(), golden["kat_mag_err"].copy()
    got = shim.loglikelihood(p(t), p(flux), p(err), 1000, p(kat), p(md), p(me))
    assert abs(got / golden["kat_logL_gmag"][0] - 1) < 1e-10
    err2 = golden["clamp_err"].copy()
    t1k = wl.time_grid(1000)
    fl = golden["n1000_flux"].copy()
    md0, me0 = np.array([1000.0, 1, 1, 1, 1]), np.full(4, 1e15)
    for k in range(4):
        pk = golden["n1000_params"][k].copy()
        got = shim.loglikelihood(p(t1k), p(fl), p(err2), 1000, p(pk), p(md0), p(me0))
        assert abs(got / golden["clamp_logL"][k] - 1) < 1e-10
    assert err2.min() == 1e-5 and np.array_equal(err2, np.maximum(golden["clamp_err"], 1e-5))
    # remove_median with the odd-N index quirk
    x = np.random.default_rng(3).standard_normal(1001)
    y = x.copy()
    shim.remove_median(p(y), 0, 1001)
    assert np.array_equal(y, x - np.sort(x)[501])
    # traj
    outs = [np.empty(1000) for _ in range(5)]
    tp = golden["kat_traj_pars"].copy()
    shim.traj(p(t), p(tp), *[p(o) for o in outs], 1000)
    assert np.abs(outs[3] / golden["kat_traj_r"] - 1).max() < 1e-12


@pytest.mark.gpu
def test_concurrent_callers_are_combined_and_correct(shim, golden):
    """The reference's OpenMP rung loop calls loglikelihood from many threads at once (mcmc_wrapper2.c:383,
    488-489).  The shim combines concurrent calls into one batched device call: every caller must get the
    value a lone call gets, whatever the interleaving, including callers with another data set."""
    import threading
    shim.hb_shim_set_memo(0)  # every call reaches the device: this test is about the combining
    t = wl.time_grid(1000)
    fl = golden["n1000_flux"].copy()
    er = np.full(1000, 3e-4)
    md0, me0 = np.array([1000.0, 1, 1, 1, 1]), np.full(4, 1e15)
    P = golden["n1000_params"][:48].copy()
    t2 = wl.time_grid(777)
    fl2, er2 = fl[:777].copy(), er[:777].copy()
    want = np.array([shim.loglikelihood(p(t), p(fl), p(er), 1000, p(P[k].copy()), p(md0), p(me0)) for k in range(48)])
    want2 = np.array([shim.loglikelihood(p(t2), p(fl2), p(er2), 777, p(P[k].copy()), p(md0), p(me0)) for k in range(8)])
    got = np.full((3, 48), np.nan)
    got2 = np.full(8, np.nan)

    def worker(tid, nthreads):
        for rep in range(3):
            for k in range(tid, 48, nthreads):
                pk = P[k].copy()
                got[rep, k] = shim.loglikelihood(p(t), p(fl), p(er), 1000, p(pk), p(md0), p(me0))
        if tid < 8:  # a second data set mixed into the same stream of calls
            pk = P[tid].copy()
            got2[tid] = shim.loglikelihood(p(t2), p(fl2), p(er2), 777, p(pk), p(md0), p(me0))

    for nthreads in (2, 8, 16):
        got[:] = np.nan
        got2[:] = np.nan
        ths = [threading.Thread(target=worker, args=(i, nthreads)) for i in range(nthreads)]
        for th in ths:
            th.start()
        for th in ths:
            th.join(timeout=120)
            assert not th.is_alive(), "a caller never got its value"
        for rep in range(3):
            assert np.array_equal(got[rep], want, equal_nan=True), nthreads
        nmix = min(8, nthreads)
        assert np.array_equal(got2[:nmix], want2[:nmix], equal_nan=True), nthreads
    # the same storm with the memo on: repeated triples are answered from the table, with the same bits
    shim.hb_shim_set_memo(1)
    h0 = shim.hb_shim_memo_hits()
    got[:] = np.nan
    got2[:] = np.nan
    ths = [threading.Thread(target=worker, args=(i, 8)) for i in range(8)]
    for th in ths:
        th.start()
    for th in ths:
        th.join(timeout=120)
        assert not th.is_alive()
    for rep in range(3):
        assert np.array_equal(got[rep], want, equal_nan=True)
    assert np.array_equal(got2, want2, equal_nan=True)
    assert shim.hb_shim_memo_hits() > h0


@pytest.mark.gpu
def test_memo_checks_every_input(shim, golden):
    """A repeated (data, parameters, magnitudes) triple is answered from the memo (the reference driver repeats
    the current state of every rung at every step, mcmc_wrapper2.c:488); anything that changes -- a parameter
    bit, a magnitude, the data arrays modified IN PLACE behind the same pointers -- is evaluated afresh."""
    t = wl.time_grid(1000)
    fl = golden["n1000_flux"].copy()
    er = np.full(1000, 3e-4)
    md, me = np.array([100.0, 9.0, 1, 1, 1]), np.array([0.05, 1e15, 1e15, 1e15])
    pk = golden["n1000_params"][3].copy()

    def call():
        return shim.loglikelihood(p(t), p(fl), p(er), 1000, p(pk), p(md), p(me))

    def fresh():
        shim.hb_shim_set_memo(0)
        v = call()
        shim.hb_shim_set_memo(1)
        return v

    shim.hb_shim_set_memo(1)
    v0 = call()
    h = shim.hb_shim_memo_hits()
    assert call() == v0 and shim.hb_shim_memo_hits() == h + 1
    fl[10] += 1e-3  # same pointer, other data
    v1 = call()
    assert shim.hb_shim_memo_hits() == h + 1 and v1 != v0 and v1 == fresh()
    md[1] += 0.1  # other magnitude
    v2 = call()
    assert shim.hb_shim_memo_hits() == h + 1 and v2 != v1 and v2 == fresh()
    pk[4] = np.nextafter(pk[4], 10.0)  # one parameter bit
    v3 = call()
    assert shim.hb_shim_memo_hits() == h + 1 and v3 == fresh()
    shim.hb_shim_set_flags(0, 0)  # Gaia term off (likelihood3.h:11): remembered values are void
    v3b = call()
    assert shim.hb_shim_memo_hits() == h + 1 and v3b != v3
    shim.hb_shim_set_flags(1, 0)
    assert call() == v3 and shim.hb_shim_memo_hits() == h + 1
    er[5] = 1e-9  # clamped in place by the call (Q2), then compared as clamped
    v4 = call()
    assert er[5] == 1e-5 and call() == v4 and shim.hb_shim_memo_hits() == h + 2 and v4 == fresh()


@pytest.mark.gpu
def test_unmodified_driver_on_the_shim_walks_the_reference_chain(tmp_path):
    """End to end at link level: the UNMODIFIED mcmc_wrapper2.c (oracle/_ref/hb_mcmc_ref_shim: only its
    /scratch prefix moved, likelihood3.c replaced by libhb_likelihood3.so on the link line) against the same
    file compiled with the reference's own likelihood3.c (oracle/_ref/hb_mcmc_ref).  Same seeds, same ran2
    streams: as long as the likelihoods agree to far more digits than any accept decision needs, the two
    runs walk the same chain, so their chain files agree line by line."""
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    cpu, gpu = os.path.join(ref_dir, "hb_mcmc_ref"), os.path.join(ref_dir, "hb_mcmc_ref_shim")
    lc = os.path.join(ref_dir, "scratch", "data", "lightcurves", "folded_lightcurves", "102289966_new.txt")
    if not (os.path.exists(cpu) and os.path.exists(gpu) and os.path.exists(lc)):
        pytest.skip("oracle/_ref drivers not built (make -C oracle ref_driver ref_driver_shim)")
    build.build_lib()
    build.build_shim()
    chain_file = os.path.join(ref_dir, "scratch", "data", "chains", "chain.102289966_gmag_OMP_71.dat")
    rows = []
    for exe in (cpu, gpu):  # same run id = same seeds (mcmc_wrapper2.c:91); the second run rewrites the file
        r = subprocess.run([exe, "400", "102289966", "0.7960497", "71"], cwd=ref_dir, capture_output=True, text=True,
                           timeout=600, env=dict(os.environ, HB_SHIM_STATS="1"))
        assert "logL=" in r.stdout, r.stdout[-400:] + r.stderr[-400:]
        rows.append(np.loadtxt(chain_file, ndmin=2))
        if exe == gpu:
            m = re.search(r"(\d+) loglikelihood calls in (\d+) batches", r.stderr)
            assert m and int(m.group(1)) > 8 * int(m.group(2)), r.stderr[-300:]  # the team's calls were combined
    a, b = rows
    assert a.shape == b.shape and a.shape[0] >= 4
    assert np.allclose(a, b, rtol=1e-9, atol=1e-12), np.abs(a - b).max()

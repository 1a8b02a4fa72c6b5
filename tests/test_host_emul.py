"""CPU coverage of the device model source: hb_device.cuh compiled for the host (intrinsics
mapped to IEEE host ops, tests/host_emul/emul.cpp) and compared with the oracle.  Catches
algebra / folding / starter bugs without a GPU; the GPU tests then only add the device libm."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from hb_mcmc_b200 import workload as wl

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMUL = os.path.join(ROOT, "tests", "host_emul")
dp = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def emul():
    src = open(os.path.join(ROOT, "hb_mcmc_b200", "csrc", "hb_device.cuh")).read()
    src = src.replace("#include <cuda_runtime.h>", "")
    with open(os.path.join(EMUL, "hb_device_host.cuh"), "w") as f:
        f.write(src)
    so = os.path.join(EMUL, "libemul.so")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", so, os.path.join(EMUL, "emul.cpp")],
                   check=True, cwd=EMUL)
    L = C.CDLL(so)
    L.emul_raw.argtypes = [dp, dp, C.c_long, dp, C.c_int]
    L.emul_finish.argtypes = [dp, C.c_long, C.c_double, C.c_double, C.c_double, dp]
    L.emul_fmod_twopi.restype = C.c_double
    L.emul_fmod_twopi.argtypes = [C.c_double]
    L.emul_prologue.argtypes = [dp, dp, dp, C.c_int, C.c_int, dp]
    L.emul_sincos.argtypes = [dp, C.c_long, dp, dp]
    L.emul_sincos_tab.argtypes = [dp, C.c_long, dp, dp]
    L.emul_sincos_table.argtypes = [dp]
    L.emul_div.argtypes = [dp, dp, C.c_long, dp, dp]
    L.emul_phase_div.argtypes = [dp, dp, C.c_long, dp]
    return L


def test_lean_sincos_accuracy(emul):
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(-10, 10, 200000), rng.uniform(-1e5, 1e5, 50000), np.arange(-40, 41) * (np.pi / 4),
                        [0.0, 1e-300, -1e-9, 2e5, -3e7]])
    s, c = np.empty_like(x), np.empty_like(x)
    emul.emul_sincos(x.ctypes.data_as(dp), x.size, s.ctypes.data_as(dp), c.ctypes.data_as(dp))
    ulp_s = np.abs(s - np.sin(x)) / np.spacing(np.abs(np.sin(x)))
    ulp_c = np.abs(c - np.cos(x)) / np.spacing(np.abs(np.cos(x)))
    assert ulp_s.max() <= 2.0 and ulp_c.max() <= 2.0, (ulp_s.max(), ulp_c.max())
    assert np.mean(s == np.sin(x)) > 0.6 and np.mean(c == np.cos(x)) > 0.6


def test_table_sincos_accuracy(emul):
    """sincos_tab (the hot loop's sin/cos): table nodes correctly rounded, results within 2 ulp, exact
    relative accuracy next to the zeros of sin and cos."""
    import mpmath as mp
    mp.mp.prec = 200
    tab = np.empty(2048)
    emul.emul_sincos_table(tab.ctypes.data_as(dp))
    for k in range(1024):
        a = 2 * mp.pi * k / 1024
        ws, wc = float(mp.sin(a)), float(mp.cos(a))
        if k % 256 == 0:  # multiples of pi/2: exact 0 and +-1 by symmetry
            ws, wc = float(round(ws)), float(round(wc))
        assert tab[2 * k] == ws + 0.0 and tab[2 * k + 1] == wc + 0.0, k
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(-10, 10, 200000), rng.uniform(-1024, 1024, 50000), rng.uniform(-1e5, 1e5, 5000),
                        np.arange(-40, 41) * (np.pi / 4), [1024.0, -1024.0, 1023.999, 1024.001],
                        np.pi + rng.uniform(-1e-3, 1e-3, 2000), np.pi / 2 + rng.uniform(-1e-6, 1e-6, 2000),
                        [0.0, 1e-300, -1e-9, 2e5, -3e7]])
    s, c = np.empty_like(x), np.empty_like(x)
    emul.emul_sincos_tab(x.ctypes.data_as(dp), x.size, s.ctypes.data_as(dp), c.ctypes.data_as(dp))
    small = np.abs(x) <= 10
    err_s = np.abs(s - np.sin(x)) / np.spacing(np.maximum(np.abs(np.sin(x)), 2.0 ** -10))
    err_c = np.abs(c - np.cos(x)) / np.spacing(np.maximum(np.abs(np.cos(x)), 2.0 ** -10))
    assert err_s[small].max() <= 2.0 and err_c[small].max() <= 2.0, (err_s[small].max(), err_c[small].max())
    assert err_s.max() <= 3.0 and err_c.max() <= 3.0, (err_s.max(), err_c.max())
    assert np.mean(s == np.sin(x)) > 0.5 and np.mean(c == np.cos(x)) > 0.5


def test_fast_division_accuracy(emul):
    rng = np.random.default_rng(2)
    a = rng.uniform(-10, 10, 200000)
    b = np.concatenate([rng.uniform(0.005, 2.0, 150000), 10.0 ** rng.uniform(-3, 6, 50000)])
    q, r = np.empty_like(a), np.empty_like(a)
    emul.emul_div(a.ctypes.data_as(dp), b.ctypes.data_as(dp), a.size, q.ctypes.data_as(dp), r.ctypes.data_as(dp))
    assert (np.abs(q - a / b) / np.abs(a / b)).max() <= 2.0 ** -45  # Newton-step quotient (see div_fast)
    assert (np.abs(r - 1 / b) / np.spacing(1 / b)).max() <= 1.0


def test_phase_division_is_ieee(emul):
    rng = np.random.default_rng(3)
    n = 400000
    P = 10.0 ** rng.uniform(-2, 3, n) * 86400.0
    x = 2 * np.pi * rng.uniform(-3000, 3000, n) * 86400.0
    out = np.empty(n)
    emul.emul_phase_div(x.ctypes.data_as(dp), P.ctypes.data_as(dp), n, out.ctypes.data_as(dp))
    assert np.array_equal(out, x / P)


def raw(L, p, t, use_table=1):
    p = np.ascontiguousarray(p, dtype=np.float64)
    t = np.ascontiguousarray(t, dtype=np.float64)
    out = np.empty(t.size)
    L.emul_raw(p.ctypes.data_as(dp), t.ctypes.data_as(dp), t.size, out.ctypes.data_as(dp), use_table)
    return out


def test_fmod_exact(emul):
    rng = np.random.default_rng(0)
    y = 2 * 3.14159265358979323846
    xs = np.concatenate([rng.uniform(-3e5, 3e5, 20000), np.arange(-50, 50) * y, np.nextafter(np.arange(1, 60) * y, 0),
                         np.nextafter(np.arange(1, 60) * y, 1e9), [0.0, -0.0, 1e-300, 5e14, -7e14]])
    for x in xs:
        assert emul.emul_fmod_twopi(float(x)) == np.fmod(x, y), x
    assert np.isnan(emul.emul_fmod_twopi(float("inf")))


@pytest.mark.parametrize("truth,N,tol", [(wl.TRUTH_A, 20000, 2e-13), (wl.TRUTH_B, 50000, 1e-14)])
def test_raw_template_truths(emul, orc, truth, N, tol):
    t = wl.time_grid(N)
    _, want = orc.calc_light_curve(t, truth, raw=True)
    got = raw(emul, truth, t)
    assert np.abs(got - want).max() < tol


def test_raw_template_random_draws(emul, orc):
    for truth, emax, seed in ((wl.TRUTH_A, 0.95, 3), (wl.TRUTH_B, 0.99, 4)):
        t = wl.time_grid(3000) * 7.0 - 5.0  # negative and positive phases, several periods
        P = wl.draw_chains(64, truth, lambda P: np.zeros(len(P)), seed=seed, e_max=emax)
        worst = 0.0
        for p in P:
            _, want = orc.calc_light_curve(t, p, raw=True)
            got = raw(emul, p, t)
            assert np.array_equal(np.isnan(got), np.isnan(want))
            scale = np.maximum(np.abs(want), 1.0)
            worst = max(worst, np.nanmax(np.abs(got - want) / scale))
        assert worst < 1e-11, worst


def test_table_starter_equals_reference_starter(emul, orc):
    """Chains with e <= 0.8 start Newton from the E(M) table: same converged root, so the raw template
    agrees with the reference-starter path (and hence the oracle) to rounding noise."""
    t = wl.time_grid(4000) * 5.0 - 3.0
    P = wl.draw_chains(96, wl.TRUTH_A, lambda P: np.array([orc.roche_overflow(p) for p in P]), seed=11, e_max=0.8)
    for k, e in enumerate([0.0, 1e-9, 0.3, 0.5, 0.6, 0.7, 0.79, 0.8]):
        q = P[k].copy()
        q[3] = e
        if not orc.roche_overflow(q):
            P[k] = q
    worst = 0.0
    for p in P:
        a, b, c = raw(emul, p, t, 1), raw(emul, p, t, 0), raw(emul, p, t, 3)  # 3 = the kernel's configuration
        scale = np.maximum(np.abs(b), 1.0)
        worst = max(worst, np.nanmax(np.abs(a - b) / scale), np.nanmax(np.abs(c - b) / scale))
        _, want = orc.calc_light_curve(t, p, raw=True)
        assert np.nanmax(np.abs(a - want) / scale) < 1e-11
        assert np.nanmax(np.abs(c - want) / scale) < 1e-11
    assert worst < 2e-13, worst


def test_table_starter_eccentric_chains(emul, orc):
    """0.8 < e <= 0.99: the table serves the samples at least 0.1 rad of mean anomaly away from periastron
    (where the reference's five steps converge, tests/tools/kepler_convergence_scan.c); inside the window the
    reference's own starter and un-converged iterates are followed.  Both must reproduce the oracle."""
    t = wl.time_grid(6000) * 3.0 - 4.0
    P = wl.draw_chains(40, wl.TRUTH_B, lambda P: np.zeros(len(P)), seed=21, e_max=0.99)
    P[:, 3] = np.linspace(0.801, 0.99, len(P))
    P[:20, 2] = np.log10(np.linspace(0.7, 9.0, 20))  # short periods: many periastron passages in the data
    worst = 0.0
    for p in P:
        _, want = orc.calc_light_curve(t, p, raw=True)
        for mode in (1, 3):  # table starter with the polynomial sincos / with the table sincos (the kernel)
            got = raw(emul, p, t, mode)
            assert np.array_equal(np.isnan(got), np.isnan(want))
            worst = max(worst, np.nanmax(np.abs(got - want) / np.maximum(np.abs(want), 1.0)))
    assert worst < 1e-11, worst


def test_table_sincos_high_e(emul, orc):
    """The table sincos in the un-converged Newton tail (e up to 0.99, reference starter)."""
    t = wl.time_grid(3000) * 7.0 - 5.0
    P = wl.draw_chains(48, wl.TRUTH_B, lambda P: np.zeros(len(P)), seed=4, e_max=0.99)
    P[:, 3] = np.linspace(0.8, 0.99, len(P))
    for p in P:
        _, want = orc.calc_light_curve(t, p, raw=True)
        got = raw(emul, p, t, 2)
        assert np.array_equal(np.isnan(got), np.isnan(want))
        assert np.nanmax(np.abs(got - want) / np.maximum(np.abs(want), 1.0)) < 1e-11


def test_finish_matches_reference_order(emul, orc):
    t = wl.time_grid(1001)
    lc, u = orc.calc_light_curve(t, wl.TRUTH_A, raw=True)
    med = np.sort(u)[orc.median_rank(u.size)]
    out = np.empty_like(u)
    emul.emul_finish(u.ctypes.data_as(dp), u.size, med, wl.TRUTH_A[19], wl.TRUTH_A[20], out.ctypes.data_as(dp))
    assert np.array_equal(out, lc)


def test_prologue_flags(emul, orc, golden):
    md = np.array([1000.0, 1, 1, 1, 1])
    me = np.full(4, 1e15)
    n = emul.emul_const_size()
    i_flag, i_info = emul.emul_const_flag_index(), emul.emul_const_info_index()
    for p, flag in zip(golden["roche_params"], golden["roche_flags"]):
        cc = np.empty(n)
        p = np.ascontiguousarray(p)
        emul.emul_prologue(p.ctypes.data_as(dp), md.ctypes.data_as(dp), me.ctypes.data_as(dp), 1, 0, cc.ctypes.data_as(dp))
        assert (int(cc[i_flag]) & 1) == int(flag)
        R1, R2, T1, T2 = orc.radii_teffs(p)
        np.testing.assert_allclose(cc[i_info:i_info + 4], [R1, R2, T1, T2], rtol=1e-14)


def test_contact_sample_takes_the_limit_of_the_area_formula(emul, orc):
    """Quirk Q10 in the model pass (hb_device.cuh, eclipse_area_dev<kGuard>): the pinned chain of
    tests/golden/contact_chain_v1.json has one sample 1.2e-8 dc from the contact d = sqrt(R1^2 - R2^2) of a 33 / 0.10
    Rsun pair, where the rounding noise of h^2 decides whether asin(h/R2) is NaN.  The reference is finite there;
    the device source must be too (it was NaN before the guard), and within the amplified-ulp bound of the tail."""
    import json
    rec = json.load(open(os.path.join(ROOT, "tests", "golden", "contact_chain_v1.json")))[0]
    p = np.array([float.fromhex(v) for v in rec["params"]])
    t = wl.time_grid(rec["N"])
    _, want = orc.calc_light_curve(t, p, raw=True)
    assert np.isfinite(want).all()
    for use_table in (3, 2, 1):
        got = raw(emul, p, t, use_table)
        assert np.isfinite(got).all(), (use_table, np.nonzero(~np.isfinite(got))[0][:5])
        i = rec["sample"]
        assert abs(got[i] - want[i]) < 1e-8 and np.abs(got - want).max() < 1e-8
        assert np.abs(np.delete(got - want, i)).max() < 1e-12


def test_table_starter_is_within_one_newton_step_of_the_root(emul):
    """The E(M) starter of the hot loop (hb_device.cuh: cubic Taylor polynomial about the nearest of 769 nodes).  The
    warp-uniform exit of the Newton loop fires when a step is below 2^-27: the starter's error is that first step, so
    'one step' needs |E0 - E| < 7.45e-9.  Asserted here for every M at e <= 0.6 (the bulk of prior draws), with the
    measured bounds at e = 0.8 (two steps next to periastron) and, at e = 0.95, outside the 0.1 rad window in which the
    table is not used; odd symmetry and the mirrored upper half are covered by the sign and range of M."""
    emul.emul_table_guess.argtypes = [C.c_double, dp, C.c_long, dp]
    rng = np.random.default_rng(11)
    two_pi = 2 * 3.14159265358979323846
    m = np.concatenate([rng.uniform(-two_pi, two_pi, 200000), np.linspace(0, two_pi, 7001)[:-1], [0.0, 1e-12, -1e-12, 3.14159265358979]])
    m = m[np.abs(m) < two_pi]

    def root(e, M):
        E = M + e * np.sin(M)
        for _ in range(60):
            E = E - (E - e * np.sin(E) - M) / (1 - e * np.cos(E))
        return E

    worst = {}
    for e in (0.0, 0.1, 0.3, 0.5, 0.6, 0.7, 0.8, 0.95):
        out = np.empty_like(m)
        emul.emul_table_guess(e, m.ctypes.data_as(dp), m.size, out.ctypes.data_as(dp))
        err = np.abs(out - root(e, m))
        if e > 0.8:  # window chains: the table serves |M| >= 0.1 rad from periastron only (kTableMinM)
            am = np.abs(m)
            err = err[np.minimum(am, two_pi - am) >= 0.1]
        worst[e] = err.max()
    assert all(worst[e] < 7.45e-9 for e in (0.0, 0.1, 0.3, 0.5, 0.6)), worst
    # measured: 1.1e-8 at e = 0.7 (0.15 % of M take a second step), 1.2e-7 at 0.8 (1.2 %), 1.4e-7 at 0.95 outside the window
    assert worst[0.7] < 2e-8 and worst[0.8] < 2e-7 and worst[0.95] < 3e-7, worst

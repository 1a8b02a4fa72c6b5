"""Memory safety without compute-sanitizer (closed on the pool): libhb_b200_dbg.so is the same source built with
-DHB_DEBUG_BOUNDS, where every indexed access of the likelihood path (candidate lists, template keys, E(M) and sin/cos
tables, select buffers, histogram bins, partial sums, padded data arrays, history rings) is asserted against its
capacity and TRAPS.  The edge-size sweep, the select stress cases, the shared-chain spreads and a sampler run go
through it and must (a) not trap and (b) give the normal build's bits; a self-test proves the assertions fire."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

import hb_mcmc_b200 as hb
from hb_mcmc_b200 import build, lib as hblib, workload as wl

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def dbg():
    path = build.build_debug_lib()
    L = hblib.load_library(path)
    c = hb.Context.__new__(hb.Context)
    h = C.c_void_p()
    assert L.hb_create(C.byref(h), 0) == 0
    c._L, c._h, c.device, c.n_points = L, h, 0, 0
    yield c
    c.close()


def test_edge_sizes_and_spreads_under_bounds_assertions(ctx, dbg):
    rng = np.random.default_rng(5)
    P = wl.draw_chains(8, wl.TRUTH_A, ctx.roche_overflow, seed=21)
    P[1, 3] = 0.9
    for N in (1, 2, 31, 33, 255, 256, 257, 1023, 1024, 1025, 2047, 2049, 4107, 4108, 4109, 7700, 12001, 20000, 50001):
        t = np.sort(rng.uniform(0, 30, N))
        flux = 1 + 1e-3 * rng.standard_normal(N)
        err = rng.uniform(1e-4, 1e-3, N)
        ctx.set_data(t, flux, err)
        dbg.set_data(t, flux, err)
        want = ctx.loglikelihood(P)
        for parts in (1, 4, 64):
            dbg.set_max_parts(parts)
            for k in (1, len(P)):
                assert np.array_equal(dbg.loglikelihood(P[:k]), want[:k], equal_nan=True), (N, parts, k)
        if N in (257, 1025, 4108, 20000):
            assert np.array_equal(dbg.light_curves(P[:2]), ctx.light_curves(P[:2]), equal_nan=True)
    # forced fallbacks: every chain misses its bracket / overflows its candidate list / re-runs for the sincos range
    dbg.set_max_parts(64)
    for sigma, rng_lim in ((0.0, 1024.0), (40.0, 1024.0), (2.5, 0.5)):
        dbg.set_bracket_sigma(sigma)
        dbg.set_sincos_range(rng_lim)
        got = dbg.loglikelihood(P)
        ref = np.abs(got - want) / np.abs(want)
        assert np.nanmax(ref) < 1e-12, (sigma, rng_lim)
    dbg.set_bracket_sigma(2.5)
    dbg.set_sincos_range(1024.0)


def test_select_and_sampler_under_bounds_assertions(ctx, dbg):
    rng = np.random.default_rng(2)
    for n in (1, 2, 255, 256, 257, 5000, 70001):
        for x in (rng.standard_normal(n), np.round(rng.standard_normal(n), 1), np.zeros(n), np.arange(n, dtype=float)[::-1].copy()):
            for k in sorted({0, n // 2, n - 1}):
                assert dbg.order_statistic(x, k) == np.sort(x)[k]
    from hb_mcmc_b200.pt import PTSampler
    t, flux, err = wl.make_dataset(3000, wl.TRUTH_A, ctx.calc_light_curve)
    ctx.set_data(t, flux, err)
    dbg.set_data(t, flux, err)
    a = PTSampler(ctx, 12, 2, float(wl.TRUTH_A[2]), seed=4, npast=8)
    b = PTSampler(dbg, 12, 2, float(wl.TRUTH_A[2]), seed=4, npast=8)
    for s in (a, b):
        s.init_random()
        s.step(30)  # past npast: DE proposals read the history rings
    for u, v in zip(a.state(), b.state()):
        assert np.array_equal(u, v, equal_nan=True)
    a.close()
    b.close()


def test_the_assertions_fire():
    """A deliberate violation in a subprocess (the trap poisons the CUDA context): index 5 against capacity 4."""
    code = (
        "import ctypes as C, sys\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "from hb_mcmc_b200 import build\n"
        "L = C.CDLL(build.build_debug_lib())\n"
        "h = C.c_void_p(); assert L.hb_create(C.byref(h), 0) == 0\n"
        "L.hb_debug_bounds_selftest.argtypes = [C.c_void_p, C.c_int]\n"
        "print('in range ->', L.hb_debug_bounds_selftest(h, 3), flush=True)\n"
        "print('out of range ->', L.hb_debug_bounds_selftest(h, 5), flush=True)\n")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert "in range -> 0" in r.stdout, r.stdout + r.stderr
    assert "out of range -> 2" in r.stdout and "bounds violation: site 99 index 5 capacity 4" in r.stdout, r.stdout + r.stderr

"""C-ABI behaviour at the edges: state errors, empty batches, non-finite parameters, re-used contexts."""
import numpy as np
import pytest

import hb_mcmc_b200 as hb
from hb_mcmc_b200 import workload as wl

pytestmark = pytest.mark.gpu


def test_calls_before_data_fail_loudly():
    c = hb.Context(0)
    try:
        with pytest.raises(hb.HBError):
            c.loglikelihood(wl.TRUTH_A[None])
        with pytest.raises(hb.HBError):
            c.light_curves(wl.TRUTH_A[None])
        # entry points that carry their own time grid need no data set
        lc = c.calc_light_curve(wl.time_grid(100), wl.TRUTH_A)
        assert lc.shape == (100,) and np.isfinite(lc).all()
        with pytest.raises((hb.HBError, ValueError)):
            c.set_data(np.zeros(3), np.zeros(4), np.zeros(3))  # ragged arrays
        with pytest.raises(hb.HBError):
            c.set_bracket_sigma(-1.0)
        with pytest.raises((hb.HBError, ValueError)):
            c.loglikelihood(np.zeros((2, 20)))  # wrong parameter count
    finally:
        c.close()


def test_non_finite_parameters_give_nan_not_garbage(ctx, orc):
    t, flux, err = wl.make_dataset(3000, wl.TRUTH_A, ctx.calc_light_curve)
    ctx.set_data(t, flux, err)
    # a NaN in any one parameter: same verdict as the reference (NaN logL -- the sampler rejects such
    # proposals, mcmc_wrapper2.c:495,505 -- or a finite value where the parameter is unused)
    P = np.tile(wl.TRUTH_A, (22, 1))
    for i in range(21):
        P[1 + i, i] = np.nan
    got = ctx.loglikelihood(P)
    want = orc.loglikelihood_batch(t, flux, err, P)
    assert np.isfinite(got[0])
    assert np.array_equal(np.isnan(got), np.isnan(want)), np.flatnonzero(np.isnan(got) != np.isnan(want))
    fin = np.isfinite(want)
    assert np.allclose(got[fin], want[fin], rtol=1e-10)
    # infinities (unreachable through the prior box; some are degenerate-but-finite in the reference):
    # no crash, and a poisoned chain does not disturb its neighbours in the batch
    Q = np.tile(wl.TRUTH_A, (43, 1))
    k = 1
    for i in range(21):
        for v in (np.inf, -np.inf):
            Q[k, i] = v
            k += 1
    gq = ctx.loglikelihood(Q)
    assert gq[0] == got[0] and gq.shape == (43,)
    clean = ctx.loglikelihood(np.tile(wl.TRUTH_A, (4, 1)))
    assert np.all(clean == got[0])


def test_huge_times_take_the_library_fmod(ctx, orc):
    """Times so large that the mean anomaly leaves the range of the in-line fmod (|M| >= 1e15): the hot pass does not
    call the library from inside its sample loop -- it flags the chain, which is evaluated again by the general pass
    with the library's exact fmod (likelihood3.c:153).  Same value as the reference, hot pass (N = 20 000, shared and
    unshared) or small-N path."""
    for N in (20000, 600):
        t = 3.0e15 + 2.0 * np.arange(N)  # days; ulp(3e15) = 0.5
        rng = np.random.default_rng(8)
        flux = 1 + 1e-3 * rng.standard_normal(N)
        err = np.full(N, 5e-4)
        ctx.set_data(t, flux, err)
        P = wl.draw_chains(700 if N == 20000 else 8, wl.TRUTH_A, lambda P: ctx.roche_overflow(P), seed=6)
        want = orc.loglikelihood_batch(t, flux, err, P[:8])
        for k in (8, len(P)):  # a shared batch, and (N = 20 000) one that fills the grid
            got = ctx.loglikelihood(P[:k])
            assert np.array_equal(np.isnan(got[:8]), np.isnan(want))
            fin = np.isfinite(want)
            assert np.allclose(got[:8][fin], want[fin], rtol=1e-10), (N, k)


def test_context_reuse_across_data_sets(ctx, orc):
    """Buffers are re-sized and re-padded on every hb_set_data: alternate long and short light curves."""
    rng = np.random.default_rng(3)
    P = wl.draw_chains(16, wl.TRUTH_A, lambda P: ctx.roche_overflow(P), seed=2)
    for N in (9000, 40, 20000, 700, 9000):
        t = np.sort(rng.uniform(0, 25, N))
        flux = 1 + 1e-3 * rng.standard_normal(N)
        err = np.full(N, 5e-4)
        ctx.set_data(t, flux, err)
        got = ctx.loglikelihood(P)
        want = orc.loglikelihood_batch(t, flux, err, P)
        assert np.allclose(got, want, rtol=1e-10), N


def test_pinned_and_pageable_host_buffers_agree(ctx):
    """hb_loglikelihood_batch reads and writes page-locked caller buffers in place over the bus when both sides are
    page-locked (k_prologue / k_chain_eval on mapped host memory, no copy), DMAs a page-locked input when only that side
    is, and stages pageable ones in chunks: the same bits every way, below and above the small-batch limit."""
    import torch
    t, flux, err = wl.make_dataset(2000, wl.TRUTH_A, ctx.calc_light_curve)
    ctx.set_data(t, flux, err)
    for n in (1, 7, 1023, 1024, 3001):
        P = wl.draw_chains(n, wl.TRUTH_A, lambda P: ctx.roche_overflow(P), seed=n)
        want = ctx.loglikelihood(P)  # pageable in, pageable out
        Pp = torch.from_numpy(P).clone().pin_memory()
        op = torch.empty(n, dtype=torch.float64).pin_memory()
        ctx.loglikelihood_into(Pp.numpy(), op.numpy())  # pinned in, pinned out
        assert np.array_equal(op.numpy(), want)
        out = np.empty(n)
        ctx.loglikelihood_into(Pp.numpy(), out)  # pinned in, pageable out
        assert np.array_equal(out, want)

"""Exact order statistic on the device (replaces the quicksort of likelihood3.c:36-105)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def cases():
    rng = np.random.default_rng(42)
    yield "random", rng.standard_normal(20000)
    yield "sorted", np.sort(rng.standard_normal(20001))
    yield "reversed", np.sort(rng.standard_normal(50000))[::-1].copy()
    yield "constant", np.full(20000, 0.97)
    yield "two_values", np.where(rng.random(30000) < 0.5, 1.0, 1.0 + 1e-16 * 4)
    yield "heavy_ties", np.round(rng.standard_normal(40000), 1)
    yield "signed_zero", np.concatenate([np.zeros(500), -np.zeros(500), rng.standard_normal(9000) * 1e-3])
    yield "plateau_eclipse", np.concatenate([np.full(15000, 1.0) + 1e-12 * rng.standard_normal(15000), np.full(5000, 0.8)])
    yield "wide_range", np.concatenate([10.0 ** rng.uniform(-300, 300, 5000), -(10.0 ** rng.uniform(-300, 300, 5000))])
    yield "inf", np.concatenate([rng.standard_normal(3000), [np.inf, -np.inf, np.inf]])
    yield "large", rng.standard_normal(300000)
    for n in (1, 2, 3, 17, 511, 512, 513, 1025):
        yield f"small{n}", rng.standard_normal(n)


@pytest.mark.parametrize("name,x", list(cases()), ids=[c[0] for c in cases()])
def test_order_statistic(ctx, name, x):
    s = np.sort(x)
    n = x.size
    ks = sorted({0, n - 1, n // 2, min(n - 1, n // 2 + 1), n // 3, (2 * n) // 3})
    for k in ks:
        got = ctx.order_statistic(x, k)
        assert got == s[k], (name, k, got, s[k])


def test_nan_input(ctx):
    x = np.random.default_rng(0).standard_normal(5000)
    x[1234] = np.nan
    assert np.isnan(ctx.order_statistic(x, 2500))


@pytest.mark.parametrize("N", [5000, 20001])
def test_missed_bracket_reruns_and_agrees(ctx, N):
    """The fused median brackets the reference's rank from a 256-sample pre-sample; a chain whose bracket
    misses is evaluated again with its template stored and selected exactly.  Forcing every chain down
    that path (sigma = 0) and widening the bracket until candidates overflow to global scratch (sigma =
    40) must give the bits of the default path."""
    from hb_mcmc_b200 import workload as wl
    t, flux, err = wl.make_dataset(N, wl.TRUTH_A, ctx.calc_light_curve)
    ctx.set_data(t, flux, err)
    P = wl.draw_chains(96, wl.TRUTH_A, lambda P: ctx.roche_overflow(P), seed=3)
    base = ctx.loglikelihood(P)
    try:
        for sigma in (0.0, 0.3, 40.0):
            ctx.set_bracket_sigma(sigma)
            assert np.array_equal(ctx.loglikelihood(P), base, equal_nan=True), sigma
    finally:
        ctx.set_bracket_sigma(2.5)
    assert np.isfinite(base).all()

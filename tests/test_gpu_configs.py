"""Every BASELINE.json configuration at its full size through the C ABI, a seeded subset of each against the
compiled reference (oracle/_ref; the oracle's restatement -- bit-identical to it, tests/test_oracle_vs_ref.py --
when the .so is absent).  Gate: |logL_gpu - logL_ref| <= 1e-10 |logL_ref|, NaN <-> NaN (likelihood3.c:809-873)."""
import numpy as np
import pytest

from conftest import rel_err
from hb_mcmc_b200 import workload as wl

pytestmark = pytest.mark.gpu
TOL = 1e-10


@pytest.fixture(scope="module")
def checker(orc):
    import oracle
    return oracle.Reference() if oracle.have_reference() else orc


def _check(got, want):
    assert np.array_equal(np.isnan(got), np.isnan(want)), "NaN pattern differs"
    fin = ~np.isnan(want)
    r = rel_err(got[fin], want[fin])
    assert r.max() <= TOL, f"max rel err {r.max():.3e}"
    return float(r.max())


def test_c2_512_chains_against_the_reference(ctx, checker):
    """C2: 4096 chains x 20 000 points; 512 of them checked."""
    N, n = 20000, 4096
    t, flux, err = wl.make_dataset(N, wl.TRUTH_A, ctx.calc_light_curve)
    ctx.set_data(t, flux, err)
    ctx.set_mags([1000, 1, 1, 1, 1], [1e15] * 4, 1, 0)
    P = wl.draw_chains(n, wl.TRUTH_A, ctx.roche_overflow, seed=1)
    P[0] = wl.TRUTH_A
    got = ctx.loglikelihood(P)
    assert np.isfinite(got).all()
    pick = np.sort(np.random.default_rng(2).choice(n, 512, replace=False))
    pick[0] = 0
    _check(got[pick], checker.loglikelihood_batch(t, flux, err, P[pick]))


def test_c4_50k_points_with_the_gaia_term(ctx, checker):
    """C4: 8192 chains x 50 000 points with the Gaia G-magnitude term (likelihood3.c:834-860); 96 chains checked,
    with and without the colour terms."""
    N, n = 50000, 8192
    t, flux, err = wl.make_dataset(N, wl.TRUTH_A, ctx.calc_light_curve)
    ctx.set_data(t, flux, err)
    G = ctx.chain_info(wl.TRUTH_A[None], 100.0)[0, 4]
    md, me = np.array([100.0, G + 0.02, 1.0, 1.0, 1.0]), np.array([0.05, 1e15, 1e15, 1e15])
    ctx.set_mags(md, me, 1, 0)
    P = wl.draw_chains(n, wl.TRUTH_A, ctx.roche_overflow, seed=4)
    P[0] = wl.TRUTH_A
    got = ctx.loglikelihood(P)
    assert np.isfinite(got).all()
    pick = np.sort(np.random.default_rng(3).choice(n, 96, replace=False))
    pick[0] = 0
    want = checker.loglikelihood_batch(t, flux, err, P[pick], md, me)
    _check(got[pick], want)
    # the Gaia term is really in there: without it the truth's logL is higher by (0.02 / 0.05)^2 / 2
    ctx.set_mags([1000, 1, 1, 1, 1], [1e15] * 4, 1, 0)
    plain = ctx.loglikelihood(P[:1])[0]
    assert abs((plain - got[0]) - 0.5 * (0.02 / 0.05) ** 2) < 1e-6
    _check(np.array([plain]), checker.loglikelihood_batch(t, flux, err, P[:1]))


def test_c5_share_200k_points_high_eccentricity(ctx, checker):
    """C5's per-GPU share: 2048 chains x 200 000 points, truth B (e = 0.95), chain eccentricities up to 0.95 -- the
    un-converged tail of the reference's five Newton steps (likelihood3.c:152-160) is inside; 64 chains checked,
    the most eccentric ones among them."""
    N, n = 200000, 2048
    t, flux, err = wl.make_dataset(N, wl.TRUTH_B, ctx.calc_light_curve)
    ctx.set_data(t, flux, err)
    ctx.set_mags([1000, 1, 1, 1, 1], [1e15] * 4, 1, 0)
    P = wl.draw_chains(n, wl.TRUTH_B, ctx.roche_overflow, seed=5, e_max=0.95)
    P[0] = wl.TRUTH_B
    got = ctx.loglikelihood(P)
    assert not np.isnan(got).any()
    by_e = np.argsort(-P[:, 3])
    pick = np.unique(np.concatenate([[0], by_e[:24], np.random.default_rng(4).choice(n, 40, replace=False)]))
    assert P[pick, 3].max() > 0.93
    _check(got[pick], checker.loglikelihood_batch(t, flux, err, P[pick]))
    # the light curve is 3.2 MB of keys per resident CTA: a second data set of a different size must still work
    t2, f2, e2 = wl.make_dataset(20000, wl.TRUTH_B, ctx.calc_light_curve)
    ctx.set_data(t2, f2, e2)
    _check(ctx.loglikelihood(P[pick[:8]]), checker.loglikelihood_batch(t2, f2, e2, P[pick[:8]]))


def test_c1_one_chain_and_c3_share(ctx, checker):
    """C1 (one chain x 20 000 points, the reference's own test_likelihoods.c case) and C3's per-GPU share
    (64 temperatures x 32 ensembles = 2048 walkers x 20 000 points): the same chains give the same bits whether
    they arrive alone, in a small batch or in the full batch."""
    N = 20000
    t, flux, err = wl.make_dataset(N, wl.TRUTH_A, ctx.calc_light_curve)
    ctx.set_data(t, flux, err)
    ctx.set_mags([1000, 1, 1, 1, 1], [1e15] * 4, 1, 0)
    P = wl.draw_chains(2048, wl.TRUTH_A, ctx.roche_overflow, seed=3)
    P[0] = wl.TRUTH_A
    full = ctx.loglikelihood(P)
    _check(full[:64], checker.loglikelihood_batch(t, flux, err, P[:64]))
    for k in (1, 2, 7, 50, 64, 300):
        assert np.array_equal(ctx.loglikelihood(P[:k]), full[:k]), k
    one = np.array([ctx.loglikelihood(P[i:i + 1])[0] for i in (5, 900, 2047)])
    assert np.array_equal(one, full[[5, 900, 2047]])

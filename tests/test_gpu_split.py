"""Batches smaller than the grid: k_chain_eval spreads every light curve over several CTAs (time-axis parts).  The
chi^2 is summed per segment of the time axis and in segment order whatever the spread is, so a chain's logL must not
depend on how many CTAs shared it, on the batch it arrived in, or on which fallback (missed bracket, candidate
overflow, sincos range) finished it.  The reference loops serially over the samples (likelihood3.c:147,649,822)."""
import numpy as np
import pytest

from conftest import rel_err
from hb_mcmc_b200 import workload as wl

pytestmark = pytest.mark.gpu


@pytest.fixture()
def restore(ctx):
    yield
    ctx.set_max_parts(64)
    ctx.set_bracket_sigma(2.5)
    ctx.set_sincos_range(1024.0)


@pytest.mark.parametrize("N,truth,emax", [(20000, "A", 0.95), (50001, "A", 0.9), (200000, "B", 0.95), (4609, "A", 0.95),
                                          (1025, "A", 0.9)])
def test_bits_do_not_depend_on_the_spread(ctx, orc, restore, N, truth, emax):
    tv = wl.TRUTH_A if truth == "A" else wl.TRUTH_B
    t, flux, err = wl.make_dataset(N, tv, ctx.calc_light_curve)
    ctx.set_data(t, flux, err)
    ctx.set_mags([1000, 1, 1, 1, 1], [1e15] * 4, 1, 0)
    P = wl.draw_chains(40, tv, ctx.roche_overflow, seed=9, e_max=emax)
    P[0] = tv
    P[1, 3] = 0.85  # an eccentric chain: the deferred range check variant of the pass
    P = P[ctx.roche_overflow(P) == 0]
    ctx.set_max_parts(1)
    base = ctx.loglikelihood(P)  # one CTA per chain
    want = orc.loglikelihood_batch(t, flux, err, P[:6])
    assert rel_err(base[:6], want).max() <= 1e-10
    for parts in (2, 4, 8, 16, 32, 64):
        ctx.set_max_parts(parts)
        for k in (1, 3, len(P)):  # the spread actually chosen also depends on the batch size
            got = ctx.loglikelihood(P[:k])
            assert np.array_equal(got, base[:k]), (N, parts, k, np.abs(got - base[:k]).max())
    # a batch that fills the grid never splits: same bits again
    big = np.vstack([P] * 20)[:700]
    got = ctx.loglikelihood(big)
    assert np.array_equal(got, np.vstack([base[:, None]] * 20)[:700, 0])


def test_fallbacks_of_a_shared_chain_give_the_same_bits(ctx, restore):
    """Forced misses of the bracket (sigma 0), overflowing candidate lists (sigma 40) and out-of-range Newton iterates
    (sincos range 0.5) send the chain that several CTAs shared to a second, whole evaluation by the CTA that arrived
    last: same bits as the undisturbed evaluation, spread or not."""
    N = 20000
    t, flux, err = wl.make_dataset(N, wl.TRUTH_A, ctx.calc_light_curve)
    ctx.set_data(t, flux, err)
    P = wl.draw_chains(24, wl.TRUTH_A, ctx.roche_overflow, seed=4)
    ctx.set_max_parts(1)
    base = ctx.loglikelihood(P)
    for parts in (1, 8, 16):
        ctx.set_max_parts(parts)
        for sigma in (0.0, 40.0, 2.5):
            ctx.set_bracket_sigma(sigma)
            assert np.array_equal(ctx.loglikelihood(P), base), (parts, sigma)
        ctx.set_bracket_sigma(2.5)
        ctx.set_sincos_range(0.5)
        got = ctx.loglikelihood(P)
        ctx.set_sincos_range(1024.0)
        # the second evaluation uses the library sincos where the table's range was (artificially) left: close, and
        # identical between spreads
        assert rel_err(got, base).max() < 1e-12
        if parts == 1:
            redo = got
        else:
            assert np.array_equal(got, redo), parts


def test_nan_roche_and_counts_in_shared_chains(ctx, golden, restore):
    """Early-outs (Roche overflow, e >= 1) and NaN templates inside a spread batch; evaluated-chain counter."""
    N = 20000
    t = wl.time_grid(N)
    ctx.set_data(t, golden["n20000_flux"], np.full(N, wl.SIGMA))
    P = np.vstack([golden["n20000_params"][:6], golden["roche_params"][:3], golden["nan_params"][:3]])
    ctx.set_max_parts(1)
    base = ctx.loglikelihood(P)
    ctx.evaluated_chains(reset=True)
    ctx.set_max_parts(16)
    got = ctx.loglikelihood(P)
    assert np.array_equal(got, base, equal_nan=True)
    n_eval = ctx.evaluated_chains(reset=True)
    early = int(np.sum(ctx.roche_overflow(P) == 1) + np.sum(~(P[:, 3] < 1.0) & (ctx.roche_overflow(P) == 0)))
    assert n_eval == len(P) - early, (n_eval, early)
    # a shared batch spreads the grid over the chains it really evaluates (the kernel lists them from their flags, in
    # two half-lists of 256): early-outs scattered through a batch that reaches into the second half
    big = np.vstack([P] * 25)[:290][np.random.default_rng(5).permutation(290)]
    ctx.set_max_parts(1)
    base_big = ctx.loglikelihood(big)
    for parts in (2, 64):
        ctx.set_max_parts(parts)
        assert np.array_equal(ctx.loglikelihood(big), base_big, equal_nan=True), parts
        assert np.array_equal(ctx.loglikelihood(big[:37]), base_big[:37], equal_nan=True), parts
    with pytest.raises(Exception):
        ctx.set_max_parts(3)

"""Oracle against the compiled, unmodified reference (oracle/_ref), bit for bit."""
import numpy as np

from hb_mcmc_b200 import workload as wl


def test_bit_identical_random_draws(orc, ref):
    rng = np.random.default_rng(123)
    for N in (777, 1000):
        t, fl, er = wl.make_dataset(N, wl.TRUTH_A, ref.calc_light_curve, seed=N)
        P = wl.draw_chains(96, wl.TRUTH_A, lambda P: np.zeros(len(P)), seed=int(rng.integers(1 << 30)))
        a = orc.loglikelihood_batch(t, fl, er, P)
        b = ref.loglikelihood_batch(t, fl, er, P)
        assert np.array_equal(a, b, equal_nan=True)
        for p in P[:6]:
            assert np.array_equal(orc.calc_light_curve(t, p), ref.calc_light_curve(t, p), equal_nan=True)
            assert orc.roche_overflow(p) == ref.roche_overflow(p)
            assert np.array_equal(orc.calc_mags(p, 250.0), ref.calc_mags(p, 250.0))
            assert orc.radii_teffs(p) == ref.radii_teffs(p)


def test_bit_identical_high_e(orc, ref):
    t, fl, er = wl.make_dataset(4000, wl.TRUTH_B, ref.calc_light_curve)
    P = wl.draw_chains(24, wl.TRUTH_B, lambda P: np.zeros(len(P)), seed=9, e_max=0.99)
    P[:, 3] = np.linspace(0.85, 0.995, len(P))
    assert np.array_equal(orc.loglikelihood_batch(t, fl, er, P), ref.loglikelihood_batch(t, fl, er, P), equal_nan=True)


def test_limits_and_sigmas(orc, ref):
    for a, b in zip(orc.set_limits(3.7), ref.set_limits(3.7)):
        assert np.array_equal(a, b)
    assert np.array_equal(orc.proposal_sigmas(1, 0), ref.proposal_sigmas())

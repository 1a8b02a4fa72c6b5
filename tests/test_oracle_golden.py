"""Oracle (C restatement) against the golden vectors generated from the compiled reference.

Pins the checker itself.  Same image / same glibc on the GPU box, so
equality is expected to be bit-exact; RTOL leaves room for a different libm only.
"""
import numpy as np

from hb_mcmc_b200 import workload as wl

RTOL = 1e-13
MSUN, SEC_DAY, RSUN = 1.9885e33, 86400.0, 6.955e10


def close(a, b, rtol=RTOL):
    np.testing.assert_allclose(a, b, rtol=rtol, atol=0, equal_nan=True)


def test_kat_traj(orc, golden):
    tr = orc.traj(golden["kat_times"], golden["kat_traj_pars"])
    for k in ("d", "Z1", "Z2", "r", "nu"):
        close(tr[k], golden["kat_traj_" + k])
    # first line of the reference's trajectories.txt
    assert abs(tr["r"][0] - 773302380561.112305) < 1.0
    assert abs(tr["nu"][0] - 2.064106) < 1e-6


def test_kat_light_curve(orc, golden):
    lc = orc.calc_light_curve(golden["kat_times"], golden["kat_params"])
    close(lc, golden["kat_lc"])
    assert np.array_equal(lc, golden["kat_lc"]) or np.allclose(lc, golden["kat_lc"], rtol=RTOL, atol=0)
    # probe values of SURVEY.md Appendix C
    assert abs(lc[0] - 0.99861996762079885) < 1e-15
    assert abs(lc.min() - 0.83752953969100619) < 1e-15
    assert abs(lc.max() - 1.0084573678173403) < 1e-15


def test_chain_scalars(orc, golden):
    p = golden["kat_params"]
    close(orc.radii_teffs(p), golden["kat_radii_teffs"])
    close(orc.calc_mags(p, 100.0), golden["kat_mags_D100"])
    assert orc.roche_overflow(p) == int(golden["kat_roche"][0]) == 0
    close([orc.alpha_beam(x) for x in golden["alpha_beam_logT"]], golden["alpha_beam"])
    lm = golden["logM_grid"]
    close([orc.getT(x) for x in lm], golden["getT"])
    close([orc.getR(x) for x in lm], golden["getR"])
    close([orc.envelope_radius(x) for x in lm], golden["envelope_radius"])
    close([orc.envelope_temp(x) for x in lm], golden["envelope_temp"])


def test_eclipse_regions(orc, golden):
    for R1, R2, d, area in golden["eclipse_cases"]:
        got = orc.eclipse_area(R1, R2, d)
        assert (np.isnan(got) and np.isnan(area)) or abs(got - area) <= RTOL * max(abs(area), 1e-300), (R1, R2, d)
    assert abs(orc.eclipse_area(1, 0.5, 0.8 * RSUN) - 0.54910621859670772) < 1e-15
    assert abs(orc.eclipse_area(1, 0.5, 1.2 * RSUN) - 0.17009800104552417) < 1e-15


def test_flux_terms(orc, golden):
    p = golden["kat_params"]
    M1, M2, Pd = 10 ** p[0], 10 ** p[1], 10 ** p[2]
    for nu, b, e, r in golden["flux_terms"]:
        close(orc.beaming(Pd, M1, M2, p[3], p[4], p[5], nu, 0.8), b)
        close(orc.ellipsoidal(Pd, M1, M2, p[3], p[4], p[5], nu, 0.83, 7.0, p[9], p[10]), e)
        close(orc.reflection(Pd, M1, M2, p[3], p[4], p[5], nu, 2.02, p[13]), r)


def test_kat_loglikelihood(orc, golden):
    t, p = golden["kat_times"], golden["kat_params"]
    flux, err = np.ones(1000), np.full(1000, 1e-3)
    md, me = golden["kat_mag_data"], golden["kat_mag_err"]
    close(orc.loglikelihood(t, flux, err, p), golden["kat_logL_nogaia"][0])
    close(orc.loglikelihood(t, flux, err, p, md, me, 1, 0), golden["kat_logL_gmag"][0])
    close(orc.loglikelihood(t, flux, err, p, md, me, 1, 1), golden["kat_logL_gmag_color"][0])
    assert abs(golden["kat_logL_nogaia"][0] - (-561350.17109085585)) < 1e-6
    assert abs(golden["kat_logL_gmag"][0] - (-562181.44455504941)) < 1e-6


def test_random_draws(orc, golden):
    md, me = golden["kat_mag_data"], golden["kat_mag_err"]
    for tag, N in (("n1000", 1000), ("n1001", 1001), ("n20000", 20000)):
        t = wl.time_grid(N)
        err = np.full(N, wl.SIGMA)
        P = golden[f"{tag}_params"]
        if N == 20000:
            P = P[:12]
        close(orc.loglikelihood_batch(t, golden[f"{tag}_flux"], err, P), golden[f"{tag}_logL"][: len(P)])
        close(orc.loglikelihood_batch(t, golden[f"{tag}_flux"], err, P, md, me), golden[f"{tag}_logL_gmag"][: len(P)])
    t = wl.time_grid(1001)
    for p, lc in zip(golden["n1001_params"][:4], golden["n1001_lc"]):
        close(orc.calc_light_curve(t, p), lc)


def test_median_rank_quirk(orc):
    # likelihood3.c:97-101: even N -> N/2, odd N -> N/2 + 1
    assert orc.median_rank(1000) == 500
    assert orc.median_rank(1001) == 501
    assert orc.median_rank(20000) == 10000


def test_high_e_nan_roche_clamp(orc, golden):
    t = wl.time_grid(20000)
    err = np.full(20000, wl.SIGMA)
    close(orc.loglikelihood_batch(t, golden["highe_flux"], err, golden["highe_params"][:9]), golden["highe_logL"][:9])
    t = wl.time_grid(1000)
    err = np.full(1000, wl.SIGMA)
    flux = golden["n1000_flux"]
    got = orc.loglikelihood_batch(t, flux, err, golden["nan_params"])
    # e == 1 exactly: 1 - e == 0 makes RocheOverflow true, and the override wins over the NaN chi^2
    want = golden["nan_logL"]
    assert np.array_equal(np.isnan(want), [False, True, True, True, False, True])
    assert np.array_equal(got, want, equal_nan=True)
    got = orc.loglikelihood_batch(t, flux, err, golden["roche_params"])
    close(got, golden["roche_logL"])
    flagged = golden["roche_flags"] == 1
    assert flagged.sum() > 5 and np.all(got[flagged] == -5e14)
    assert [orc.roche_overflow(p) for p in golden["roche_params"]] == list(golden["roche_flags"].astype(int))
    close(orc.loglikelihood_batch(t, flux, golden["clamp_err"], golden["n1000_params"][:16]), golden["clamp_logL"])


def test_sampler_pieces(orc, golden):
    lo, hi, ml, mh, g = orc.set_limits(2.0)
    assert np.array_equal(np.stack([lo, hi, ml, mh, g.astype(np.float64)]), golden["limits_P2"])
    assert np.array_equal(orc.proposal_sigmas(1, 0), golden["sigmas"])
    close([orc.get_logP(p, g) for p in golden["logP_params"]], golden["logP"])
    # quirk Q4: e is not bounded above (mode 0.99 is neither reflect nor periodic)
    y = golden["kat_params"].copy()
    y[3] = 1.3
    y[5] = 4.0   # periodic wrap
    y[4] = -0.2  # reflect
    out = orc.enforce_bounds(y, lo, hi, ml, mh, np.log10(2.0))
    assert out[3] == 1.3 and abs(out[5] - (4.0 - 2 * np.pi)) < 1e-12 and abs(out[4] - 0.2) < 1e-15
    assert out[2] == np.log10(2.0)
    # swap rule (mcmc_wrapper2.c:796-816)
    temp = 1.4 ** np.arange(4)
    acc, idx = orc.pt_swap_pair([0, 1, 2, 3], temp, [-10.0, -5.0, -7.0, -1.0], 0, 0.5)
    assert acc == 1 and list(idx) == [1, 0, 2, 3]
    acc, idx = orc.pt_swap_pair([0, 1, 2, 3], temp, [-5.0, -50.0, -7.0, -1.0], 0, 0.5)
    assert acc == 0 and list(idx) == [0, 1, 2, 3]

"""Generate tests/golden/golden_v1.npz from the UNMODIFIED reference compiled in oracle/_ref.

Run in the build container (where /root/reference exists):

    make -C oracle ref && python tests/golden/make_golden.py

The reference ships no expected outputs, so the golden vectors are
outputs of the reference's own code (`likelihood3.c`, gcc -O3 -std=c99, glibc 2.39) on the one
reference-supplied input (test_likelihoods.c:33-36) plus seeded random prior draws.  Every
array is float64 and stored bit-exactly.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from hb_mcmc_b200 import workload as wl  # noqa: E402

MSUN, SEC_DAY, RSUN = 1.9885e33, 86400.0, 6.955e10


def draws(n, truth, ref, seed, e_max=0.95, reject_roche=True):
    fn = (lambda P: np.array([ref.roche_overflow(p) for p in P])) if reject_roche else (lambda P: np.zeros(len(P)))
    return wl.draw_chains(n, truth, fn, seed=seed, e_max=e_max)


def main():
    R = oracle.Reference()
    Rc = oracle.Reference(color=True)
    S = oracle.ReferenceSampler()
    G = {}
    kat = wl.TRUTH_A
    G["kat_params"] = kat

    # (1) traj for the KAT vector on the reference's own grid (test_likelihoods.c:27-30,52-61)
    t_kat = 4 * (np.arange(1000) / 1000)
    tp = np.array([10 ** kat[0] * MSUN, 10 ** kat[1] * MSUN, 10 ** kat[2] * SEC_DAY, kat[3], kat[4], kat[5],
                   kat[6] * SEC_DAY])
    tr = R.traj(t_kat, tp)
    G["kat_times"] = t_kat
    G["kat_traj_pars"] = tp
    for k, v in tr.items():
        G["kat_traj_" + k] = v
    # (2) full template
    G["kat_lc"] = R.calc_light_curve(t_kat, kat)
    # (3) per-chain scalars
    G["kat_radii_teffs"] = np.array(R.radii_teffs(kat))
    G["kat_mags_D100"] = R.calc_mags(kat, 100.0)
    G["kat_roche"] = np.array([R.roche_overflow(kat)])
    lt = np.array([3.4, 3.5, 3.55, 3.7, 3.8, 3.9, 4.2, 4.5, 4.7])
    G["alpha_beam_logT"] = lt
    G["alpha_beam"] = np.array([R.alpha_beam(x) for x in lt])
    lm = np.linspace(-1.6, 2.1, 75)
    G["logM_grid"] = lm
    G["getT"] = np.array([R.getT(x) for x in lm])
    G["getR"] = np.array([R.getR(x) for x in lm])
    G["envelope_radius"] = np.array([R.envelope_radius(x) for x in lm])
    G["envelope_temp"] = np.array([R.envelope_temp(x) for x in lm])
    # eclipse_area across all four regions and their boundaries (likelihood3.c:368-386)
    ecl = []
    for R1, R2 in ((1.0, 0.5), (0.5, 1.0), (2.0225368206330931, 0.83266527058320738), (1.0, 1.0)):
        big, small = max(R1, R2), min(R1, R2)
        dc = np.sqrt(big * big - small * small)
        for d in (0.0 if big != small else 0.1, 0.5 * (big - small), big - small, 0.5 * (big - small + dc), dc,
                  np.nextafter(dc, 9.0), 0.5 * (dc + big + small), 0.999999 * (big + small), big + small,
                  1.2 * (big + small), 0.8, 1.2):
            ecl.append((R1, R2, d * RSUN, R.eclipse_area(R1, R2, d * RSUN)))
    G["eclipse_cases"] = np.array(ecl)
    # flux terms at a few anomalies (likelihood3.c:224-337)
    terms = []
    M1, M2, Pd = 10 ** kat[0], 10 ** kat[1], 10 ** kat[2]
    for nu in np.linspace(-3.1, 3.1, 13):
        terms.append((nu,
                      R.beaming(Pd, M1, M2, kat[3], kat[4], kat[5], nu, 0.8),
                      R.ellipsoidal(Pd, M1, M2, kat[3], kat[4], kat[5], nu, 0.83, 7.0, kat[9], kat[10]),
                      R.reflection(Pd, M1, M2, kat[3], kat[4], kat[5], nu, 2.02, kat[13])))
    G["flux_terms"] = np.array(terms)

    # (4) loglikelihood with and without the Gaia term (Appendix C of SURVEY.md)
    flux1, err1 = np.ones(1000), np.full(1000, 1e-3)
    md = np.array([100, 4.5, 0.1, 0, -0.05])
    me = np.array([0.05, 0.1, 0.1, 0.1])
    G["kat_logL_nogaia"] = np.array([R.loglikelihood(t_kat, flux1, err1, kat)])
    G["kat_mag_data"], G["kat_mag_err"] = md, me
    G["kat_logL_gmag"] = np.array([R.loglikelihood(t_kat, flux1, err1, kat, md, me)])
    G["kat_logL_gmag_color"] = np.array([Rc.loglikelihood(t_kat, flux1, err1, kat, md, me)])

    # (5) random prior draws, Roche-rejected, at even / odd / full-size N
    for tag, N, n in (("n1000", 1000, 256), ("n1001", 1001, 128), ("n20000", 20000, 48)):
        t, fl, er = wl.make_dataset(N, kat, R.calc_light_curve)
        P = draws(n, kat, R, seed=11 + N)
        P[0] = kat
        G[f"{tag}_flux"] = fl
        G[f"{tag}_params"] = P
        G[f"{tag}_logL"] = R.loglikelihood_batch(t, fl, er, P)
        G[f"{tag}_logL_gmag"] = R.loglikelihood_batch(t, fl, er, P, md, me)
    # a few full templates at the odd size (median index quirk Q3)
    t1001 = wl.time_grid(1001)
    G["n1001_lc"] = np.stack([R.calc_light_curve(t1001, p) for p in G["n1001_params"][:4]])

    # (6) un-converged Kepler tail: e in {0.9, 0.95, 0.99}, truth B geometry (P = 39.8 d)
    tB, flB, erB = wl.make_dataset(20000, wl.TRUTH_B, R.calc_light_curve)
    PB = draws(48, wl.TRUTH_B, R, seed=5, e_max=0.99)
    PB[:, 3] = np.tile([0.9, 0.95, 0.99], 16)
    PB[0] = wl.TRUTH_B
    PB = PB[np.array([R.roche_overflow(p) for p in PB]) == 0]
    G["highe_flux"] = flB
    G["highe_params"] = PB
    G["highe_logL"] = R.loglikelihood_batch(tB, flB, erB, PB)
    G["highe_lc0"] = R.calc_light_curve(tB, wl.TRUTH_B)

    # (7) NaN cases: e >= 1 is reachable (quirk Q4)
    Pn = G["n1000_params"][:6].copy()
    Pn[:, 3] = [1.0, 1.0000001, 1.2, 2.0, 1.0, 1.5]
    t1k, fl1k, er1k = wl.make_dataset(1000, kat, R.calc_light_curve)
    G["nan_params"] = Pn
    G["nan_logL"] = R.loglikelihood_batch(t1k, fl1k, er1k, Pn)

    # (8) Roche-flagged cases (not rejected): exactly -BIG_NUM/2
    Pr = draws(64, kat, R, seed=77, reject_roche=False)
    G["roche_params"] = Pr
    G["roche_flags"] = np.array([R.roche_overflow(p) for p in Pr], dtype=np.float64)
    G["roche_logL"] = R.loglikelihood_batch(t1k, fl1k, er1k, Pr)

    # (9) sigma < 1e-5 clamp (likelihood3.c:824-827)
    er_small = er1k.copy()
    er_small[::7] = 1e-7
    er_small[3] = 0.0
    G["clamp_err"] = er_small
    G["clamp_logL"] = R.loglikelihood_batch(t1k, fl1k, er_small, G["n1000_params"][:16])

    # sampler pieces: priors (mcmc_wrapper2.c:703-765), limits, sigmas
    G["logP_params"] = G["n1000_params"][:32]
    G["logP"] = np.array([S.get_logP(p) for p in G["logP_params"]])
    lo, hi, ml, mh, gf = R.set_limits(2.0)
    G["limits_P2"] = np.stack([lo, hi, ml, mh, gf.astype(np.float64)])
    G["sigmas"] = R.proposal_sigmas()

    out = os.path.join(ROOT, "tests", "golden", "golden_v1.npz")
    np.savez_compressed(out, **G)
    print("wrote", out, f"{os.path.getsize(out) / 1024:.0f} KiB,", len(G), "arrays")
    print("kat_logL_nogaia", repr(G["kat_logL_nogaia"][0]), "(SURVEY Appendix C: -561350.17109085585)")


if __name__ == "__main__":
    main()

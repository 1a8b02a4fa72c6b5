"""The measured parity distribution, pinned.

BASELINE.json's gate is |logL_gpu - logL_ref| <= 1e-10 |logL_ref|.  A wide campaign (profiles/*parity_scan*) finds
it met by 99.99 % of chains and exceeded, by up to 5e-10, where ONE in-eclipse sample sits next to
d = sqrt(R1^2 - R2^2): there the reference's asin(h/R) area formula (likelihood3.c:372-376) amplifies a 1-ulp
difference of the projected separation by ~1e7 -- and the reference compiled with FMA contraction differs from
itself by the same amount on the same chains (tests/test_outlier_fixtures.py, CPU).  The stated bound of DESIGN.md
section 5, asserted here on a seeded scan of > 20 000 chains:  >= 99.98 % of chains within 1e-10, every chain
within 1e-9, NaN <-> NaN on every chain."""
import json
import os

import numpy as np
import pytest

import parity_scan_lib as ps
from conftest import rel_err

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def checker(orc):
    import oracle
    return oracle.Reference() if oracle.have_reference() else orc


def test_seeded_scan_meets_the_stated_bound(ctx, checker):
    # passes 8..12 of the round-1 campaign (SEED_OFFSET 5000): they hold four of its five chains above 1e-10
    rel, records = ps.scan(ctx, checker, range(8, 13), seed_offset=5000)
    assert rel.size >= 20000
    frac_ok = float(np.mean(rel <= 1e-10))
    assert frac_ok >= 0.9998, f"only {frac_ok:.5f} of {rel.size} chains within 1e-10"
    assert rel.max() <= 1e-9, f"worst chain {rel.max():.3e}"
    assert np.median(rel) < 5e-14


def test_pinned_outlier_chains(ctx, checker):
    """The chains of the campaign that exceeded 1e-10, kept as fixtures: still finite, still within the amended
    bound, and the stored reference value is the reference's (the fixture did not rot)."""
    from hb_mcmc_b200 import workload as wl
    path = os.path.join(HERE, "golden", "outliers_v1.json")
    recs = json.load(open(path))
    assert len(recs) >= 4
    for key in sorted({(r["truth"], r["N"]) for r in recs}):
        group = [r for r in recs if (r["truth"], r["N"]) == key]
        truth = ps.TRUTHS[key[0]]
        t, fl, er = wl.make_dataset(key[1], truth, checker.calc_light_curve)
        P = np.array([[float.fromhex(v) for v in r["params"]] for r in group])
        want = np.array([float.fromhex(r["logL_ref"]) for r in group])
        assert np.array_equal(checker.loglikelihood_batch(t, fl, er, P), want)
        ctx.set_data(t, fl, er)
        got = ctx.loglikelihood(P)
        assert np.isfinite(got).all()
        assert rel_err(got, want).max() <= 1e-9


def test_contact_chain_is_finite(ctx, checker):
    """Quirk Q10 in the model pass: one sample of this pinned chain (set B, N = 20 000, seed 5902, chain 309) sits
    1.2e-8 dc from the contact d = sqrt(R1^2 - R2^2), where the rounding noise of h^2 decides whether asin(h/R2) is
    NaN.  The reference is finite; the kernel takes the formula's limit there (eclipse_area_dev<kGuard>) instead of
    rolling the same dice with other bits -- it answered NaN on this chain before the guard."""
    from hb_mcmc_b200 import workload as wl
    rec = json.load(open(os.path.join(HERE, "golden", "contact_chain_v1.json")))[0]
    truth = ps.TRUTHS[rec["truth"]]
    t, fl, er = wl.make_dataset(rec["N"], truth, checker.calc_light_curve)
    P = np.array([[float.fromhex(v) for v in rec["params"]]])
    want = float.fromhex(rec["logL_ref"])
    assert checker.loglikelihood_batch(t, fl, er, P)[0] == want
    ctx.set_data(t, fl, er)
    got = ctx.loglikelihood(P)[0]
    assert np.isfinite(got) and abs(got - want) <= 1e-9 * abs(want), (got, want)
    lc = ctx.light_curves(P)[0]
    assert np.isfinite(lc).all()
    assert abs(lc[rec["sample"]] - float.fromhex(rec["lc_ref_at_sample"])) < 1e-8

"""Statistical parity of the device sampler with the UNMODIFIED reference driver on a real folded
light curve (TIC 102289966, 375 points).  tests/golden/pt_reference_runs.json holds four runs of
the reference binary (20 000 iterations, 50 rungs; see make_pt_golden.py).  The RNG streams differ
by design, so the comparison is distributional: 32 independent device ladders against the
reference runs for (a) the driver's own acceptance statistics and (b) how fast the cold rung
climbs in log-likelihood."""
import json
import os

import numpy as np
import pytest

from hb_mcmc_b200.pt import PTSampler

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_device_sampler_matches_reference_driver_statistics(ctx):
    ref = json.load(open(os.path.join(GOLD, "pt_reference_runs.json")))
    with open(os.path.join(GOLD, f"lc_{ref['tic']}_new.txt")) as f:
        n = int(f.readline())
        d = np.loadtxt(f)
    assert d.shape == (n, 3)
    ctx.set_data(d[:, 0], d[:, 1], d[:, 2])
    ctx.set_mags([1000, 1, 1, 1, 1], [1e15] * 4, 1, 0)  # no magnitude file for this TIC (mcmc_wrapper2.c:319-328)
    E = 32
    s = PTSampler(ctx, ref["nchains"], E, ref["log10_period"], seed=2024, npast=ref["npast"], quirks=True)
    s.init_random()
    checkpoints = [1000, 5000, 10000, 19900]
    traj = {}
    done = 0
    for cp in checkpoints:
        s.step(cp + 1 - done)  # the reference logs after iteration index cp
        done = cp + 1
        traj[cp] = s.cold()[1].copy()
    cnt = s.counters()
    it = cnt["iterations"].astype(float)
    acc = cnt["acc_slot0"] / it                       # the reference's `acc` (chain id 0 accepted)
    deacc = cnt["de_acc_slot0"] / np.maximum(cnt["de_trials_slot0"], 1)
    ref_acc = np.mean([np.mean(r["acc_prints"]) for r in ref["runs"]])
    ref_de = np.mean([np.mean(r["deacc_prints"]) for r in ref["runs"]])
    # (a) acceptance of Gaussian+DE proposals and the (as-compiled, nearly useless) DE jumps
    assert abs(acc.mean() - ref_acc) < 0.1, (acc.mean(), ref_acc)
    assert deacc.mean() < 0.03 and ref_de < 0.03, (deacc.mean(), ref_de)
    # (b) the cold rung's climb: the reference runs sit inside the spread of the device ladders
    for cp in checkpoints:
        i = ref["runs"][0]["iter"].index(cp)
        r = np.array([-run["cold_logL"][i] for run in ref["runs"]])
        g = -traj[cp]
        assert np.all(np.isfinite(g))
        lo, hi = np.quantile(np.log10(g), [0.02, 0.98])
        med = np.median(np.log10(r))
        assert lo - 0.35 <= med <= hi + 0.35, (cp, 10 ** lo, 10 ** hi, r)
    # and both improve by orders of magnitude over the random start
    assert np.median(-traj[19900]) < 0.2 * np.median(-traj[1000])
    s.close()

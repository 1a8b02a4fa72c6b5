"""Device Gaia-colour sampler (hb_gaia_pt_*) against the oracle's restatement of GAIA_mcmc.c fed the
same Philox streams, plus launch-splitting / logging invariants."""
import ctypes as C

import numpy as np
import pytest

import oracle as orc_mod
from hb_mcmc_b200.gaia import GaiaSampler

pytestmark = pytest.mark.gpu
dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)

# TIC 186260283 and TIC 293950421 of the reference's data/color_mag/cp_data_4-21-2022.csv
STARS_D = np.array([234.296, 1173.39])
STARS = np.array([[7.16094512, -0.0066265000000005, 0.0212387299999997, -0.0066558899999999],
                  [11.64856136, 0.0606254999999993, -0.1073223099999987, 0.1480928299999992]])
STARS_E = np.array([[0.0230834782584296, 0.0367165032930016, 0.0586013200752551, 0.0086725406204692],
                    [0.0305112273197476, 0.1983198992966981, 0.06026411193706, 0.0095166038816294]])


def p(a):
    return a.ctypes.data_as(dp)


@pytest.fixture(scope="module")
def L(orc):
    return orc_mod.gaia_protos(orc.lib)


def test_step_matches_oracle(ctx, orc, L):
    T, E, npast, seed = 9, 4, 6, 777
    s = GaiaSampler(ctx, n_ens=E, n_temps=T, seed=seed, npast=npast)
    D = STARS_D[np.arange(E) % 2]
    data, err = STARS[np.arange(E) % 2], STARS_E[np.arange(E) % 2]
    s.set_data(D, data, err)
    s.init_random()
    lo, hi, ml, mh, g = orc_mod.gaia_limits(L)
    gp = g.ctypes.data_as(ip)
    sigma = np.array([1e-2, 1e-2, 0., 0., 0., 0.])
    temps = s.temps
    x, logL, index = s.state()
    assert np.all(x >= lo) and np.all(x <= hi) and np.array_equal(index, np.tile(np.arange(T), (E, 1)))
    for w in range(E * T):
        want = orc.gaia_model_likelihood(data[w // T], err[w // T], x[w], D[w // T])
        assert np.isclose(logL[w], want, rtol=1e-11, atol=1e-9), (w, logL[w], want)
    history = np.zeros((E * T, npast, 6))
    n_de = n_acc = n_swap = 0
    for it in range(80):
        s.run(1, log=False)
        y_gpu, logLy, logPy, jump = s.proposal()
        x_new, logL_new, index_new = s.state()
        index_exp = index.copy()
        for ens in range(E):
            for j in range(T):
                r = ens * T + j
                c = ens * T + index[ens, j]
                y = np.empty(6)
                lp = C.c_double()
                xc = np.ascontiguousarray(x[c])
                jt = L.orc_gaia_propose(seed, r, it, temps[j], npast, p(xc), p(np.ascontiguousarray(history[r])), p(lo),
                                        p(hi), p(ml), p(mh), gp, p(sigma), p(y), C.byref(lp))
                assert jt == jump[r], (it, r)
                n_de += jt == 2
                assert np.allclose(y_gpu[r], y, rtol=1e-12, atol=1e-14), (it, r, jt, y_gpu[r], y)
                assert np.isclose(logPy[r], lp.value, rtol=1e-12, atol=1e-13)
                want = orc.gaia_model_likelihood(data[ens], err[ens], y_gpu[r], D[ens])
                assert np.isclose(logLy[r], want, rtol=1e-11, atol=1e-9), (it, r, logLy[r], want)
                # decision with the device's own likelihood / prior values
                acc = L.orc_gaia_accept(seed, r, it, temps[j], logL[c], logLy[r],
                                        L.orc_gaia_get_logP(p(xc), p(lo), p(hi), gp), logPy[r])
                n_acc += acc
                assert np.array_equal(x_new[c], y_gpu[r] if acc else x[c]), (it, r, acc)
                assert logL_new[c] == (logLy[r] if acc else logL[c])
            idx_e = np.ascontiguousarray(index_exp[ens], dtype=np.int32)
            fill = np.empty(T, dtype=np.int32)
            ll_e = np.ascontiguousarray(logL_new[ens * T:(ens + 1) * T])
            n_swap += L.orc_gaia_swap_ensemble(seed, ens, it, T, p(temps), idx_e.ctypes.data_as(ip), p(ll_e),
                                               fill.ctypes.data_as(ip))
            index_exp[ens] = idx_e
            for j in range(T):
                history[ens * T + j, it % npast] = x_new[ens * T + fill[j]]
        assert np.array_equal(index_new, index_exp), it
        x, logL, index = x_new, logL_new, index_new
    assert np.array_equal(s.history(), history)
    assert n_de > 100 and n_acc > 30 and n_swap > 30, (n_de, n_acc, n_swap)
    cnt = s.counters()
    assert cnt["iterations"].tolist() == [80] * E and cnt["proposed"].tolist() == [80 * T] * E
    assert int(cnt["accepted"].sum()) == n_acc and int(cnt["swaps_accepted"].sum()) == n_swap
    xm, lm = s.map()
    cold = np.array([logL[e * T + index[e, 0]] for e in range(E)])
    assert np.all(lm >= cold)
    s.close()


def test_one_launch_equals_many(ctx):
    """A run is a pure function of (seed, state, iteration): 1 x 500 iterations == 7 uneven launches,
    and the thinned log is the concatenation of the pieces."""
    kw = dict(n_ens=5, n_temps=20, seed=31)
    a, b = GaiaSampler(ctx, **kw), GaiaSampler(ctx, **kw)
    for s in (a, b):
        s.set_data(STARS_D[0], STARS[0], STARS_E[0])
        s.init_random()
    chain_a, rung_a = a.run(500, thin=10)
    parts = [b.run(n, thin=10) for n in (1, 9, 10, 95, 185, 0, 200)]
    chain_b = np.concatenate([c for c, _ in parts], axis=1)
    rung_b = np.concatenate([r for _, r in parts], axis=1)
    assert chain_a.shape == (5, 50, 7) and rung_a.shape == (5, 50, 20)
    assert np.array_equal(chain_a, chain_b) and np.array_equal(rung_a, rung_b)
    for u, v in zip(a.state(), b.state()):
        assert np.array_equal(u, v)
    assert np.array_equal(a.history(), b.history())
    assert a.iteration == b.iteration == 500
    # the last record is the cold rung after iteration 490; logL column 0 of the rung log is the same number
    assert np.array_equal(chain_a[:, :, 0], rung_a[:, :, 0])
    # ensembles are independent replicas: same star, different streams
    assert len({chain_a[e, -1, 0] for e in range(5)}) == 5
    a.close()
    b.close()


def test_errors(ctx):
    from hb_mcmc_b200 import HBError
    with pytest.raises(HBError):
        GaiaSampler(ctx, n_temps=33)
    s = GaiaSampler(ctx, n_ens=2)
    with pytest.raises(HBError):
        s.init_random()  # no data yet
    s.set_data(STARS_D[0], STARS[0], STARS_E[0])
    with pytest.raises(HBError):
        s.run(10)  # no state yet
    s.init_random()
    chain, rung = s.run(0)
    assert chain.shape == (2, 0, 7)
    s.close()


def test_statistical_parity_with_reference_runs(ctx):
    """Posterior summaries of the cold chain against 8 runs of the UNMODIFIED GAIA_mcmc.c
    (tests/golden/gaia_reference_runs.json, made by tests/golden/make_gaia_golden.py): same star, same
    chain length, same thinning and burn-in; tolerances are a few times the run-to-run scatter of the
    reference itself."""
    import json
    import os
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, "golden"))
    from make_gaia_golden import summarise
    ref = json.load(open(os.path.join(here, "golden", "gaia_reference_runs.json")))
    E, niter = 16, 400000
    s = GaiaSampler(ctx, n_ens=E, n_temps=ref["nchains"], seed=2024, npast=ref["npast"])
    s.set_data(ref["distance"], STARS[0], STARS_E[0])
    s.init_random()
    chain, rung = s.run(niter, thin=ref["thin"])
    assert chain.shape == (E, niter // ref["thin"], 7)
    mine = [summarise(chain[e], rung[e]) for e in range(E)]
    med = lambda runs, key, i=None: float(np.median([r[key] if i is None else r[key][i] for r in runs]))
    sd = lambda runs, key, i=None: float(np.std([r[key] if i is None else r[key][i] for r in runs]))
    checks = [("cold_logL_mean", None, 0.25), ("m_hi_q", 1, 0.02), ("m_lo_q", 1, 0.006), ("rr_hi_q", 1, 0.12),
              ("rr_lo_q", 1, 0.6), ("at_hi_q", 1, 0.25), ("at_lo_q", 1, 0.6), ("m_hi_q", 0, 0.03), ("m_hi_q", 2, 0.02),
              ("cold_logL_q", 1, 0.25)]
    for key, i, tol in checks:
        a, b = med(mine, key, i), med(ref["runs"], key, i)
        assert abs(a - b) < tol, (key, i, a, b, sd(mine, key, i), sd(ref["runs"], key, i))
    ra = np.median([m["rung_logL_mean"] for m in mine], axis=0)
    rb = np.median([r["rung_logL_mean"] for r in ref["runs"]], axis=0)
    assert np.allclose(ra, rb, rtol=0.04, atol=0.1), (ra, rb)
    cnt = s.counters()
    acc = cnt["accepted"].sum() / cnt["proposed"].sum()
    assert 0.05 < acc < 0.8, acc
    s.close()

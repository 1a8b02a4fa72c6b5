"""Oracle restatement of the Gaia-colour sampler (GAIA_mcmc.c) against the unmodified file compiled
with oracle/gsl_stub (GSL supplies random numbers only; the tests feed the draws), bit for bit."""
import ctypes as C
import math
import os

import numpy as np
import pytest

import oracle as orc_mod

HERE = os.path.dirname(os.path.abspath(__file__))
dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)


def p(a):
    return a.ctypes.data_as(dp)


@pytest.fixture(scope="module")
def gref():
    try:
        return orc_mod.ReferenceGaia()
    except FileNotFoundError:
        pytest.skip("oracle/_ref/libref_gaia.so not built (reference sources absent)")


@pytest.fixture(scope="module")
def L(orc):
    return orc_mod.gaia_protos(orc.lib)


# TIC 186260283 of the reference's data/color_mag/cp_data_4-21-2022.csv (dist, Gmag0, BmV0, VmG0, GmT0 and errors)
STAR_D = 234.296
STAR = np.array([7.16094512, -0.0066265000000005, 0.0212387299999997, -0.0066558899999999])
STAR_E = np.array([0.0230834782584296, 0.0367165032930016, 0.0586013200752551, 0.0086725406204692])


def test_scalar_pieces(L, orc, gref):
    rng = np.random.default_rng(5)
    for a, b in zip(orc_mod.gaia_limits(L), gref.set_limits()):
        assert np.array_equal(a, b)
    s = gref.proposal_sigmas()
    assert s[0] == 1e-2 and s[1] == 1e-2 and np.isnan(s[2:]).all()  # init_proposals sets two of six
    lo, hi, ml, mh, g = orc_mod.gaia_limits(L)
    for _ in range(200):
        x = lo + rng.random(6) * (hi - lo)
        assert L.orc_gaia_get_logP(p(x), p(lo), p(hi), g.ctypes.data_as(ip)) == gref.get_logP(x)
        assert L.orc_gaia_gaussian(x[2], 0.3, 1.7) == gref.gaussian(x[2], 0.3, 1.7)
        assert np.array_equal(orc.gaia_get_mags(x, STAR_D), gref.get_mags(x, STAR_D))
        assert orc.gaia_model_likelihood(STAR, STAR_E, x, STAR_D) == gref.model_likelihood(STAR, STAR_E, x, STAR_D)


def stream(L, seed, rid, it, stage, n=96):
    u = np.empty(n)
    L.orc_pt_uniforms(seed, rid, it, stage, n, p(u))
    return u


def box_muller(u1, u2):
    r = math.sqrt(-2.0 * math.log(u1))
    a = 6.283185307179586 * u2
    return r * math.cos(a), r * math.sin(a)


def split_draws(u, it, npast):
    """Walk the proposal stream the way orc_gaia_propose does and split it into the uniforms and the
    normals GAIA_mcmc.c would ask GSL for (worst case: DE scaling normals + Gaussian fallback)."""
    k = 2
    uni, nor = [u[0], u[1]], []
    de = (u[1] < 0.5) and (it > npast)

    def six(k):
        z = []
        for _ in range(3):
            z.extend(box_muller(u[k], u[k + 1]))
            k += 2
        return z, k
    if de:
        a = int(u[k] * npast); uni.append(u[k]); k += 1
        b = a
        while b == a:
            b = int(u[k] * npast); uni.append(u[k]); k += 1
        uni.append(u[k]); k += 1
        if uni[-1] < 0.9:
            z, k = six(k)
            nor += z
    z, k = six(k)  # Gaussian jump, or the fallback of a short DE jump (left over when not taken)
    nor += z
    return uni, nor, de


def test_step_bit_identical_to_reference(L, gref):
    """Propose + accept + swaps + history of the oracle == run_chain / ptmcmc of GAIA_mcmc.c when both
    consume the same draws, over enough iterations to fill the DE history and exercise every branch."""
    T, NP, seed = gref.NCHAINS, gref.NPAST, 4242
    lo, hi, ml, mh, g = orc_mod.gaia_limits(L)
    gp = g.ctypes.data_as(ip)
    sigma = np.array([1e-2, 1e-2, 0., 0., 0., 0.])
    temp = 1.2 ** np.arange(T)
    temp[0] = 1.0
    for i in range(1, T):
        temp[i] = temp[i - 1] * 1.2
    rng = np.random.default_rng(11)
    x_ref = lo + rng.random((T, 6)) * (hi - lo)
    x_orc = x_ref.copy()
    hist_ref = np.zeros((T, NP, 6))
    hist_orc = np.zeros((T, NP, 6))
    index_ref = np.arange(T, dtype=np.int32)
    index_orc = index_ref.copy()
    ll = np.array([gref.model_likelihood(STAR, STAR_E, x_ref[i], STAR_D) for i in range(T)])
    logL_ref, logL_orc = ll.copy(), ll.copy()
    libc = C.CDLL(None)
    n_de = n_acc = n_swap = n_fallback = 0
    for it in range(260):
        for j in range(T):
            u = stream(L, seed, j, it, 0)
            uni, nor, de = split_draws(u, it, NP)
            ua = stream(L, seed, j, it, 1, 2)[0]
            slot = int(index_orc[j])
            y = np.empty(6)
            lp = C.c_double()
            xs = x_orc[slot].copy()
            jt = L.orc_gaia_propose(seed, j, it, temp[j], NP, p(xs), p(np.ascontiguousarray(hist_orc[j])), p(lo), p(hi),
                                    p(ml), p(mh), gp, p(sigma), p(y), C.byref(lp))
            logLy = gref.model_likelihood(STAR, STAR_E, y, STAR_D)
            acc = L.orc_gaia_accept(seed, j, it, temp[j], logL_orc[slot], logLy,
                                    L.orc_gaia_get_logP(p(xs), p(lo), p(hi), gp), lp.value)
            if acc:
                x_orc[slot] = y
                logL_orc[slot] = logLy
            n_de += jt == 2
            n_fallback += de and jt == 1
            n_acc += acc
            y_ref, left = gref.run_chain(it, x_ref, sigma, temp, index_ref, hist_ref, j, STAR, STAR_E, STAR_D, logL_ref,
                                         uni + [ua], nor)
            assert np.array_equal(y_ref, y), (it, j, jt)
            assert left[0] == 0, (it, j, left)
        assert np.array_equal(x_ref, x_orc) and np.array_equal(logL_ref, logL_orc), it
        # swaps + history (run_mcmc, GAIA_mcmc.c:741-748): rand() drives the reference, the same
        # numbers are handed to the oracle's pair rule
        libc.srand(1000 + it)
        rr = [libc.rand() for _ in range(2 * T)]
        libc.srand(1000 + it)
        RAND_MAX = 2147483647
        k = it - (it // NP) * NP
        for c in range(T):
            gref.ptmcmc(index_ref, temp, logL_ref)
            hist_ref[c, k] = x_ref[index_ref[c]]
            b = int(float(rr[2 * c]) / RAND_MAX * float(T - 1))
            beta = float(rr[2 * c + 1]) / RAND_MAX
            n_swap += L.orc_pt_swap_pair(index_orc.ctypes.data_as(ip), p(temp), p(logL_orc), b, beta)
            hist_orc[c, k] = x_orc[index_orc[c]]
        assert np.array_equal(index_ref, index_orc), it
    assert n_de > 500 and n_acc > 300 and n_swap > 300, (n_de, n_acc, n_swap)


def test_swap_ensemble_fill_order(L):
    """orc_gaia_swap_ensemble = T x (swap proposal, then record the slot rung k holds)."""
    T, seed = 7, 9
    temp = 1.2 ** np.arange(T)
    rng = np.random.default_rng(2)
    for it in range(50):
        logL = -rng.random(T) * 30
        idx = rng.permutation(T).astype(np.int32)
        idx2 = idx.copy()
        fill = np.empty(T, dtype=np.int32)
        n = L.orc_gaia_swap_ensemble(seed, 3, it, T, p(temp), idx.ctypes.data_as(ip), p(logL), fill.ctypes.data_as(ip))
        u = stream(L, seed, 0x80000000 | 3, it, 2, 2 * T)
        m = 0
        for k in range(T):
            b = min(int(u[2 * k] * (T - 1)), T - 2)
            m += L.orc_pt_swap_pair(idx2.ctypes.data_as(ip), p(temp), p(logL), b, u[2 * k + 1])
            assert fill[k] == idx2[k]
        assert n == m and np.array_equal(idx, idx2)

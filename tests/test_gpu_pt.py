"""Device-resident PT step against the oracle's restatement fed the same Philox stream."""
import ctypes as C

import numpy as np
import pytest

import hb_mcmc_b200 as hb
from hb_mcmc_b200 import workload as wl
from hb_mcmc_b200.pt import PTSampler

pytestmark = pytest.mark.gpu
dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)


def p(a):
    return a.ctypes.data_as(dp)


@pytest.fixture(scope="module")
def setup(ctx, orc):
    N = 600
    t, flux, err = wl.make_dataset(N, wl.TRUTH_A, ctx.calc_light_curve)
    ctx.set_data(t, flux, err)
    ctx.set_mags([1000, 1, 1, 1, 1], [1e15] * 4, 1, 0)
    return t, flux, err


@pytest.mark.parametrize("quirks", [True, False])
def test_step_matches_oracle(ctx, orc, setup, quirks):
    t, flux, err = setup
    T, E, npast, seed = 6, 5, 8, 12345
    logP = float(wl.TRUTH_A[2])
    s = PTSampler(ctx, T, E, logP, seed=seed, npast=npast, quirks=quirks)
    s.init_random()
    lo, hi, ml, mh, gauss = orc.set_limits(10.0 ** logP)
    sigma = orc.proposal_sigmas(1, 0)
    temps = s.temps
    L = orc.lib
    L.orc_pt_propose.restype = C.c_int
    L.orc_pt_propose.argtypes = [C.c_ulonglong, C.c_uint, C.c_uint, C.c_double, C.c_int, C.c_int, dp, dp, dp, dp, dp, dp,
                                 ip, dp, C.c_double, dp, dp]
    L.orc_pt_accept.restype = C.c_int
    L.orc_pt_accept.argtypes = [C.c_ulonglong, C.c_uint, C.c_uint] + [C.c_double] * 5
    L.orc_pt_swap_ensemble.restype = C.c_int
    L.orc_pt_swap_ensemble.argtypes = [C.c_ulonglong, C.c_uint, C.c_uint, C.c_int, dp, ip, dp]

    x, logL, index = s.state()
    # initial state: inside the prior box, period pinned, logL equals a direct evaluation
    assert np.all(x >= lo - 1e-12) and np.all(x <= hi + 1e-12) and np.all(x[:, 2] == logP)
    assert np.array_equal(logL, ctx.loglikelihood(x), equal_nan=True)
    history = np.zeros((E * T, npast, 21))
    n_de = n_acc = n_swap = 0
    for it in range(30):
        s.step(1)
        y_gpu, logLy, logPy = s.proposal()
        x_new, logL_new, index_new = s.state()
        index_exp = index.copy()
        for ens in range(E):
            for j in range(T):
                r = ens * T + j
                c = ens * T + index[ens, j]
                y = np.empty(21)
                lp = C.c_double()
                xc = np.ascontiguousarray(x[c])
                hr = np.ascontiguousarray(history[r])
                jt = L.orc_pt_propose(seed, r, it, temps[j], npast, int(quirks), p(xc), p(hr), p(lo), p(hi), p(ml), p(mh),
                                      gauss.ctypes.data_as(ip), p(sigma), logP, p(y), C.byref(lp))
                n_de += jt == 2
                # DE jumps "as compiled" are ~4000 x a history difference.  The reference (and the oracle)
                # bounce them back one reflection at a time, each with its own rounding; the device removes
                # whole multiples of the box in one step, so the two agree to ~n_bounces x ulp(jump)
                atol = 1e-13 if jt == 1 else 1e-7
                ok = np.isclose(y_gpu[c], y, rtol=1e-11, atol=atol, equal_nan=True)
                assert ok.all(), (it, r, jt, y_gpu[c][~ok], y[~ok])
                assert np.isclose(logPy[c], lp.value, rtol=1e-11 if jt == 1 else 1e-5) or (np.isnan(logPy[c]) and np.isnan(lp.value))
                # accept decision with the GPU's own likelihood values
                acc = L.orc_pt_accept(seed, r, it, temps[j], logL[c], logLy[c], orc.get_logP(xc, gauss), logPy[c])
                n_acc += acc
                want_x = y_gpu[c] if acc else x[c]
                assert np.array_equal(x_new[c], want_x, equal_nan=True), (it, r, acc)
                assert (logL_new[c] == (logLy[c] if acc else logL[c])) or np.isnan(logL_new[c])
                history[r, it % npast] = x_new[c]
            idx_e = np.ascontiguousarray(index_exp[ens], dtype=np.int32)
            ll_e = np.ascontiguousarray(logL_new[ens * T:(ens + 1) * T])
            n_swap += L.orc_pt_swap_ensemble(seed, ens, it, T, p(temps), idx_e.ctypes.data_as(ip), p(ll_e))
            index_exp[ens] = idx_e
        assert np.array_equal(index_new, index_exp), it
        # the proposals' likelihood is the batched likelihood of y
        assert np.array_equal(logLy, ctx.loglikelihood(y_gpu), equal_nan=True)
        x, logL, index = x_new, logL_new, index_new
    assert n_de > 50 and n_acc > 10 and n_swap > 5
    cnt = s.counters()
    assert cnt["iterations"].tolist() == [30] * E and cnt["proposed"].tolist() == [30 * T] * E
    assert int(cnt["accepted"].sum()) == n_acc and int(cnt["swaps_accepted"].sum()) == n_swap
    xc, lc = s.cold()
    for ens in range(E):
        c0 = ens * T + index[ens, 0]
        assert np.array_equal(xc[ens], x[c0]) and lc[ens] == logL[c0]
    xm, lm = s.map()
    assert np.all(lm >= lc)
    assert np.array_equal(s.logL_by_rung(), np.take_along_axis(logL.reshape(E, T), index, axis=1))
    s.close()


def test_sampler_recovers_truth_region(ctx, setup):
    """Short statistical check: started at the truth, the cold chains stay near chi^2 ~ N and the
    acceptance rate is sane (the reference prints acc ~ 0.2-0.5 for its hand-set jump sizes)."""
    t, flux, err = setup
    T, E = 8, 16
    s = PTSampler(ctx, T, E, float(wl.TRUTH_A[2]), seed=7, npast=50, quirks=True)
    s.set_state(np.tile(wl.TRUTH_A, (T * E, 1)))
    s.step(300)
    cnt = s.counters()
    rate = cnt["accepted"].sum() / cnt["proposed"].sum()
    assert 0.02 < rate < 0.9, rate
    _, lc = s.cold()
    assert np.all(np.isfinite(lc)) and np.median(-2 * lc) < 3 * len(t)
    s.close()


def test_graph_replay_equals_single_steps(ctx, setup):
    """hb_pt_step(n >= 4) replays a captured CUDA graph of one iteration; it must walk exactly the
    same chain as n single (un-captured) iterations, including after the data set changes."""
    t, flux, err = setup
    a = PTSampler(ctx, 7, 4, float(wl.TRUTH_A[2]), seed=99, npast=12)
    b = PTSampler(ctx, 7, 4, float(wl.TRUTH_A[2]), seed=99, npast=12)
    for s in (a, b):
        s.set_one_launch(False)  # the stream-ordered kernels (light curves this short would take the one-launch loop)
    a.init_random()
    b.init_random()
    for _ in range(40):
        a.step(1)
    b.step(40)
    xa, la, ia = a.state()
    xb, lb, ib = b.state()
    assert np.array_equal(xa, xb, equal_nan=True) and np.array_equal(la, lb, equal_nan=True) and np.array_equal(ia, ib)
    assert a.iteration == b.iteration == 40
    for k, v in a.counters().items():
        assert np.array_equal(v, b.counters()[k]), k
    # new data set: buffers are re-created, the captured graph must not be reused blindly
    ctx.set_data(t[:300], flux[:300], err[:300])
    a.set_state(xa)
    b.set_state(xb)
    for _ in range(8):
        a.step(1)
    b.step(8)
    assert np.array_equal(a.state()[0], b.state()[0], equal_nan=True)
    ctx.set_data(t, flux, err)
    a.close()
    b.close()


@pytest.mark.parametrize("T,E,N", [(50, 1, 375), (7, 4, 600), (64, 9, 1024), (12, 3, 163)])
def test_one_launch_loop_equals_stream_ordered_steps(ctx, setup, T, E, N):
    """Short light curves: hb_pt_step runs the whole loop in ONE launch (k_pt_run: a CTA per walker, one grid-wide
    barrier per iteration).  Same device functions in the same order as the five stream-ordered kernels per iteration:
    states, proposals, permutations, MAP and counters must be identical bit for bit -- in uneven chunks, across a
    change of the data set, and for more walkers than fit one wave of the stream-ordered likelihood kernel."""
    t, flux, err = setup
    ctx.set_data(t[:N] if N <= len(t) else wl.time_grid(N), flux[:N] if N <= len(t) else np.resize(flux, N),
                 err[:N] if N <= len(t) else np.resize(err, N))
    a = PTSampler(ctx, T, E, float(wl.TRUTH_A[2]), seed=5, npast=10)
    b = PTSampler(ctx, T, E, float(wl.TRUTH_A[2]), seed=5, npast=10)
    b.set_one_launch(False)
    a.init_random()
    b.init_random()
    done = 0
    for chunk in (1, 2, 30, 7):  # past npast: DE proposals from the history rings
        a.step(chunk)
        b.step(chunk)
        done += chunk
        for u, v in zip(a.state() + a.proposal() + a.map(), b.state() + b.proposal() + b.map()):
            assert np.array_equal(u, v, equal_nan=True), (chunk, done)
        ca, cb = a.counters(), b.counters()
        assert all(np.array_equal(ca[k], cb[k]) for k in ca) and a.iteration == b.iteration == done
    assert ctx.evaluated_chains(reset=True) > 0
    a.close()
    b.close()
    ctx.set_data(t, flux, err)

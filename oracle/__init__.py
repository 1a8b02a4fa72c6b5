"""CPU parity checker for the HB_MCMC hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package; the product (``hb_mcmc_b200``)
never does.  Two back ends, same call surface:

* :class:`Oracle`   -- ``oracle/libhb_oracle.so``: the C restatement in ``hb_oracle.c``.
* :class:`Reference` -- ``oracle/_ref/libref_lik3.so``: the UNMODIFIED reference
  ``likelihood3.c`` compiled in the build container from ``/root/reference/src`` (the
  prebuilt ``.so`` travels to the GPU box; the sources never enter this repo).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
NPARS = 21
BIG_NUM = 1.0e15

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def _p(a):
    return a.ctypes.data_as(_dp)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def build(ref: bool = True) -> None:
    """Compile the restatement and, when /root/reference is present, the reference."""
    subprocess.run(["make", "-s", "-C", HERE, "oracle"], check=True, stdout=sys.stderr)
    if ref:
        subprocess.run(["make", "-s", "-C", HERE, "ref"], check=True, stdout=sys.stderr)
        if os.path.exists("/root/reference/src/mcmc_wrapper2.c"):
            # the unmodified driver, stand-alone (statistical goldens, the PT baseline of bench.py) and linked
            # against the product's link-level drop-in libhb_likelihood3.so when that has been built
            targets = ["ref_driver", "ref_fma"]  # ref_fma: tests/tools/ref_self_consistency.py only
            if os.path.exists(os.path.join(os.path.dirname(HERE), "hb_mcmc_b200", "csrc", "libhb_likelihood3.so")):
                targets.append("ref_driver_shim")
                if os.path.exists("/root/reference/src/pyHB.pyx"):
                    targets.append("pyhb")  # the unmodified Cython binding on the reference's C and on the drop-in
            subprocess.run(["make", "-s", "-C", HERE] + targets, check=True, stdout=sys.stderr)


def have_reference() -> bool:
    return os.path.exists(os.path.join(HERE, "_ref", "libref_lik3.so"))


DEFAULT_MAG_DATA = np.array([1000.0, 1.0, 1.0, 1.0, 1.0])  # mcmc_wrapper2.c:322-327
DEFAULT_MAG_ERR = np.array([BIG_NUM] * 4)


class Oracle:
    """ctypes front end of the C restatement (hb_oracle.c)."""

    kind = "port"

    def __init__(self):
        path = os.path.join(HERE, "libhb_oracle.so")
        if not os.path.exists(path):
            build(ref=False)
        L = self.lib = C.CDLL(path)
        for name in ("orc_getT", "orc_getR", "orc_envelope_temp", "orc_envelope_radius", "orc_alpha_beam"):
            getattr(L, name).restype = C.c_double
            getattr(L, name).argtypes = [C.c_double]
        L.orc_eclipse_area.restype = C.c_double
        L.orc_eclipse_area.argtypes = [C.c_double] * 3
        L.orc_beaming.restype = C.c_double
        L.orc_beaming.argtypes = [C.c_double] * 8
        L.orc_ellipsoidal.restype = C.c_double
        L.orc_ellipsoidal.argtypes = [C.c_double] * 11
        L.orc_reflection.restype = C.c_double
        L.orc_reflection.argtypes = [C.c_double] * 9
        L.orc_radii_teffs.argtypes = [_dp] * 5
        L.orc_traj.argtypes = [_dp] * 7 + [C.c_long]
        L.orc_calc_light_curve_ex.argtypes = [_dp, C.c_long, _dp, _dp, _dp]
        L.orc_calc_mags.argtypes = [_dp, C.c_double, _dp]
        L.orc_gaia_get_mags.argtypes = [_dp, C.c_double, _dp]
        L.orc_gaia_model_likelihood.restype = C.c_double
        L.orc_gaia_model_likelihood.argtypes = [_dp, _dp, _dp, C.c_double]
        L.orc_roche_overflow.restype = C.c_int
        L.orc_roche_overflow.argtypes = [_dp]
        L.orc_loglikelihood.restype = C.c_double
        L.orc_loglikelihood.argtypes = [_dp, _dp, _dp, C.c_long, _dp, _dp, _dp, C.c_int, C.c_int]
        L.orc_loglikelihood_batch.argtypes = [_dp, _dp, _dp, C.c_long, _dp, C.c_long, _dp, _dp, C.c_int, C.c_int, _dp]
        L.orc_set_limits.argtypes = [_dp, _dp, _dp, _dp, _ip, C.c_double]
        L.orc_proposal_sigmas.argtypes = [_dp, C.c_int, C.c_int]
        L.orc_get_logP.restype = C.c_double
        L.orc_get_logP.argtypes = [_dp, _ip]
        L.orc_enforce_bounds.argtypes = [_dp] * 5 + [C.c_double, C.c_double]
        L.orc_pt_swap_pair.restype = C.c_int
        L.orc_pt_swap_pair.argtypes = [_ip, _dp, _dp, C.c_int, C.c_double]
        L.orc_hastings.restype = C.c_double
        L.orc_hastings.argtypes = [C.c_double] * 5
        L.orc_median_rank.restype = C.c_long
        L.orc_median_rank.argtypes = [C.c_long]

    # scalar helpers
    def getT(self, x): return self.lib.orc_getT(x)
    def getR(self, x): return self.lib.orc_getR(x)
    def envelope_temp(self, x): return self.lib.orc_envelope_temp(x)
    def envelope_radius(self, x): return self.lib.orc_envelope_radius(x)
    def alpha_beam(self, x): return self.lib.orc_alpha_beam(x)
    def eclipse_area(self, R1, R2, d): return self.lib.orc_eclipse_area(R1, R2, d)
    def beaming(self, *a): return self.lib.orc_beaming(*a)
    def ellipsoidal(self, *a): return self.lib.orc_ellipsoidal(*a)
    def reflection(self, *a): return self.lib.orc_reflection(*a)
    def median_rank(self, n): return self.lib.orc_median_rank(n)

    def radii_teffs(self, pars):
        p = _f64(pars)
        o = [C.c_double() for _ in range(4)]
        self.lib.orc_radii_teffs(_p(p), *[C.byref(x) for x in o])
        return tuple(x.value for x in o)

    def traj(self, times, traj_pars):
        t = _f64(times)
        tp = _f64(traj_pars)
        out = [np.empty(t.size) for _ in range(5)]
        self.lib.orc_traj(_p(t), _p(tp), *[_p(o) for o in out], t.size)
        return dict(zip(("d", "Z1", "Z2", "r", "nu"), out))

    def calc_light_curve(self, times, pars, raw=False):
        t = _f64(times)
        p = _f64(pars)
        out = np.empty(t.size)
        rw = np.empty(t.size)
        self.lib.orc_calc_light_curve_ex(_p(t), t.size, _p(p), _p(out), _p(rw))
        return (out, rw) if raw else out

    def calc_mags(self, pars, D):
        p = _f64(pars)
        o = np.empty(4)
        self.lib.orc_calc_mags(_p(p), D, _p(o))
        return o

    def gaia_get_mags(self, p6, D):
        p = _f64(p6)
        o = np.empty(4)
        self.lib.orc_gaia_get_mags(_p(p), D, _p(o))
        return o

    def gaia_model_likelihood(self, data, err, p6, D):
        return self.lib.orc_gaia_model_likelihood(_p(_f64(data)), _p(_f64(err)), _p(_f64(p6)), D)

    def roche_overflow(self, pars):
        return self.lib.orc_roche_overflow(_p(_f64(pars)))

    def loglikelihood(self, t, flux, err, pars, mag_data=None, magerr=None, use_gmag=1, use_color=0):
        t, flux = _f64(t), _f64(flux)
        err = _f64(err).copy()  # clamped in place by the callee (quirk Q2)
        md = _f64(DEFAULT_MAG_DATA if mag_data is None else mag_data)
        me = _f64(DEFAULT_MAG_ERR if magerr is None else magerr)
        return self.lib.orc_loglikelihood(_p(t), _p(flux), _p(err), t.size, _p(_f64(pars)), _p(md), _p(me),
                                          use_gmag, use_color)

    def loglikelihood_batch(self, t, flux, err, params, mag_data=None, magerr=None, use_gmag=1, use_color=0,
                            nthreads=None):
        t, flux, err = _f64(t), _f64(flux), _f64(err)
        P = _f64(params).reshape(-1, NPARS)
        md = _f64(DEFAULT_MAG_DATA if mag_data is None else mag_data)
        me = _f64(DEFAULT_MAG_ERR if magerr is None else magerr)
        out = np.empty(P.shape[0])
        if nthreads:
            os.environ["OMP_NUM_THREADS"] = str(nthreads)
        self.lib.orc_loglikelihood_batch(_p(t), _p(flux), _p(err), t.size, _p(P), P.shape[0], _p(md), _p(me),
                                         use_gmag, use_color, _p(out))
        return out

    def set_limits(self, lc_period):
        lo, hi, ml, mh = (np.empty(NPARS) for _ in range(4))
        g = np.empty(NPARS, dtype=np.int32)
        self.lib.orc_set_limits(_p(lo), _p(hi), _p(ml), _p(mh), g.ctypes.data_as(_ip), lc_period)
        return lo, hi, ml, mh, g

    def proposal_sigmas(self, use_gmag=1, use_color=0):
        s = np.empty(NPARS)
        self.lib.orc_proposal_sigmas(_p(s), use_gmag, use_color)
        return s

    def get_logP(self, pars, gauss):
        g = np.ascontiguousarray(gauss, dtype=np.int32)
        return self.lib.orc_get_logP(_p(_f64(pars)), g.ctypes.data_as(_ip))

    def enforce_bounds(self, y, lo, hi, ml, mh, log_lc_period):
        y = _f64(y).copy()
        self.lib.orc_enforce_bounds(_p(y), _p(_f64(lo)), _p(_f64(hi)), _p(_f64(ml)), _p(_f64(mh)),
                                    log_lc_period, 10.0 ** log_lc_period)
        return y

    def pt_swap_pair(self, index, temp, logL, b, beta):
        idx = np.ascontiguousarray(index, dtype=np.int32).copy()
        acc = self.lib.orc_pt_swap_pair(idx.ctypes.data_as(_ip), _p(_f64(temp)), _p(_f64(logL)), b, beta)
        return acc, idx

    def hastings(self, logLx, logLy, logPx, logPy, temp):
        return self.lib.orc_hastings(logLx, logLy, logPx, logPy, temp)


class Reference:
    """ctypes front end of the compiled, unmodified reference (oracle/_ref)."""

    kind = "reference"

    def __init__(self, color: bool = False, variant: str = ""):
        # variant "fma": the same unmodified file compiled with FMA contraction (`make -C oracle ref_fma`), only
        # used to show how far the reference moves under a legal change of its own compilation
        name = "libref_lik3_color.so" if color else ("libref_lik3_%s.so" % variant if variant else "libref_lik3.so")
        path = os.path.join(HERE, "_ref", name)
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (run `make -C oracle ref` where /root/reference exists)")
        L = self.lib = C.CDLL(path)
        self.use_gmag, self.use_color = 1, int(color)  # likelihood3.h:11-12
        for name in ("_getT", "_getR", "envelope_Temp", "envelope_Radius", "get_alpha_beam"):
            getattr(L, name).restype = C.c_double
            getattr(L, name).argtypes = [C.c_double]
        L.eclipse_area.restype = C.c_double
        L.eclipse_area.argtypes = [C.c_double] * 3
        L.beaming.restype = C.c_double
        L.beaming.argtypes = [C.c_double] * 8
        L.ellipsoidal.restype = C.c_double
        L.ellipsoidal.argtypes = [C.c_double] * 11
        L.reflection.restype = C.c_double
        L.reflection.argtypes = [C.c_double] * 9
        L.calc_radii_and_Teffs.argtypes = [_dp] * 5
        L.traj.argtypes = [_dp] * 7 + [C.c_int]
        L.calc_mags.argtypes = [_dp, C.c_double] + [_dp] * 4
        L.RocheOverflow.restype = C.c_int
        L.RocheOverflow.argtypes = [_dp]
        L.loglikelihood.restype = C.c_double
        L.loglikelihood.argtypes = [_dp, _dp, _dp, C.c_long, _dp, _dp, _dp]
        L.ref_loglikelihood_batch.restype = C.c_int
        L.ref_loglikelihood_batch.argtypes = [_dp, _dp, _dp, C.c_long, _dp, C.c_long, _dp, _dp, C.c_int, _dp]
        L.ref_calc_light_curve.restype = C.c_int
        L.ref_calc_light_curve.argtypes = [_dp, C.c_long, _dp, _dp]
        L.set_limits.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double]
        L.initialize_proposals.argtypes = [_dp, C.c_void_p]

    def getT(self, x): return self.lib._getT(x)
    def getR(self, x): return self.lib._getR(x)
    def envelope_temp(self, x): return self.lib.envelope_Temp(x)
    def envelope_radius(self, x): return self.lib.envelope_Radius(x)
    def alpha_beam(self, x): return self.lib.get_alpha_beam(x)
    def eclipse_area(self, R1, R2, d): return self.lib.eclipse_area(R1, R2, d)
    def beaming(self, *a): return self.lib.beaming(*a)
    def ellipsoidal(self, *a): return self.lib.ellipsoidal(*a)
    def reflection(self, *a): return self.lib.reflection(*a)

    def radii_teffs(self, pars):
        p = _f64(pars).copy()
        o = [C.c_double() for _ in range(4)]
        self.lib.calc_radii_and_Teffs(_p(p), *[C.byref(x) for x in o])
        return tuple(x.value for x in o)

    def traj(self, times, traj_pars):
        t = _f64(times).copy()
        tp = _f64(traj_pars).copy()
        out = [np.empty(t.size) for _ in range(5)]
        self.lib.traj(_p(t), _p(tp), *[_p(o) for o in out], t.size)
        return dict(zip(("d", "Z1", "Z2", "r", "nu"), out))

    def calc_light_curve(self, times, pars):
        t = _f64(times).copy()
        p = _f64(pars).copy()
        out = np.empty(t.size)
        if self.lib.ref_calc_light_curve(_p(t), t.size, _p(p), _p(out)) != 0:
            raise RuntimeError("reference thread could not be started")
        return out

    def calc_mags(self, pars, D):
        p = _f64(pars).copy()
        o = [C.c_double() for _ in range(4)]
        self.lib.calc_mags(_p(p), D, *[C.byref(x) for x in o])
        return np.array([x.value for x in o])

    def roche_overflow(self, pars):
        return self.lib.RocheOverflow(_p(_f64(pars).copy()))

    def loglikelihood(self, t, flux, err, pars, mag_data=None, magerr=None, **_):
        return float(self.loglikelihood_batch(t, flux, err, np.asarray(pars)[None, :], mag_data, magerr,
                                              nthreads=1)[0])

    def loglikelihood_batch(self, t, flux, err, params, mag_data=None, magerr=None, nthreads=None, **_):
        t, flux = _f64(t).copy(), _f64(flux).copy()
        err = _f64(err).copy()
        P = _f64(params).reshape(-1, NPARS)
        md = _f64(DEFAULT_MAG_DATA if mag_data is None else mag_data).copy()
        me = _f64(DEFAULT_MAG_ERR if magerr is None else magerr).copy()
        out = np.empty(P.shape[0])
        nthreads = nthreads or os.cpu_count() or 1
        rc = self.lib.ref_loglikelihood_batch(_p(t), _p(flux), _p(err), t.size, _p(P), P.shape[0], _p(md),
                                              _p(me), int(nthreads), _p(out))
        if rc != 0:
            raise RuntimeError("reference threads could not be started")
        return out

    def set_limits(self, lc_period):
        """likelihood3.c:986 -> (lo, hi, mode_lo, mode_hi, gauss_flag) as arrays."""
        limited = (C.c_double * (2 * NPARS))()
        limits = (C.c_double * (2 * NPARS))()
        gauss = (C.c_int * NPARS)()
        self.lib.set_limits(limited, limits, gauss, lc_period)
        ld = np.array(limited).reshape(NPARS, 2)
        lm = np.array(limits).reshape(NPARS, 2)
        return lm[:, 0].copy(), lm[:, 1].copy(), ld[:, 0].copy(), ld[:, 1].copy(), np.array(gauss, dtype=np.int32)

    def proposal_sigmas(self):
        s = np.zeros(NPARS)
        self.lib.initialize_proposals(_p(s), None)
        return s


class ReferenceSampler:
    """Deterministic sampler pieces of the reference driver (oracle/_ref/libref_mcmc.so)."""

    def __init__(self):
        path = os.path.join(HERE, "_ref", "libref_mcmc.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        L = self.lib = C.CDLL(path)
        L.get_logP.restype = C.c_double
        L.get_logP.argtypes = [_dp, C.c_void_p, C.c_void_p, C.c_void_p]
        L.gaussian.restype = C.c_double
        L.gaussian.argtypes = [C.c_double] * 3
        L.set_limits.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double]

    def get_logP(self, pars, lc_period=1.0):
        limited = (C.c_double * (2 * NPARS))()
        limits = (C.c_double * (2 * NPARS))()
        gauss = (C.c_int * NPARS)()
        self.lib.set_limits(limited, limits, gauss, lc_period)
        return self.lib.get_logP(_p(_f64(pars).copy()), limited, limits, gauss)


GAIA_NPARS = 6


def gaia_protos(L):
    """argtypes of the oracle's Gaia-sampler restatement (hb_oracle.c, GAIA_mcmc.c flavour)."""
    ip = C.POINTER(C.c_int)
    L.orc_gaia_gaussian.restype = C.c_double
    L.orc_gaia_gaussian.argtypes = [C.c_double] * 3
    L.orc_gaia_get_logP.restype = C.c_double
    L.orc_gaia_get_logP.argtypes = [_dp, _dp, _dp, ip]
    L.orc_gaia_set_limits.argtypes = [_dp, _dp, _dp, _dp, ip]
    L.orc_gaia_enforce_bounds.argtypes = [_dp] * 5
    L.orc_gaia_propose.restype = C.c_int
    L.orc_gaia_propose.argtypes = [C.c_ulonglong, C.c_uint, C.c_uint, C.c_double, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp,
                                   ip, _dp, _dp, _dp]
    L.orc_gaia_accept.restype = C.c_int
    L.orc_gaia_accept.argtypes = [C.c_ulonglong, C.c_uint, C.c_uint] + [C.c_double] * 5
    L.orc_gaia_swap_ensemble.restype = C.c_int
    L.orc_gaia_swap_ensemble.argtypes = [C.c_ulonglong, C.c_uint, C.c_uint, C.c_int, _dp, ip, _dp, ip]
    L.orc_pt_uniforms.argtypes = [C.c_ulonglong, C.c_uint, C.c_uint, C.c_uint, C.c_int, _dp]
    return L


def gaia_limits(L):
    """(lo, hi, mode_lo, mode_hi, gauss) of orc_gaia_set_limits."""
    lo, hi, ml, mh = (np.empty(GAIA_NPARS) for _ in range(4))
    g = np.empty(GAIA_NPARS, dtype=np.int32)
    L.orc_gaia_set_limits(_p(lo), _p(hi), _p(ml), _p(mh), g.ctypes.data_as(C.POINTER(C.c_int)))
    return lo, hi, ml, mh, g


class ReferenceGaia:
    """The unmodified GAIA_mcmc.c (oracle/_ref/libref_gaia.so, main renamed) linked against
    oracle/gsl_stub.  `feed` queues the uniforms / normals its next gsl calls will return."""

    NCHAINS, NPAST = 20, 100

    def __init__(self):
        path = os.path.join(HERE, "_ref", "libref_gaia.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        L = self.lib = C.CDLL(path)
        vp = C.c_void_p
        L.gaussian.restype = C.c_double
        L.gaussian.argtypes = [C.c_double] * 3
        L.get_logP.restype = C.c_double
        L.get_logP.argtypes = [_dp, vp, vp, vp]
        L.get_mags.argtypes = [_dp, C.c_double, _dp]
        L.model_likelihood.restype = C.c_double
        L.model_likelihood.argtypes = [_dp, _dp, _dp, _dp, C.c_double]
        L.set_limits.argtypes = [vp, vp, vp]
        L.init_proposals.argtypes = [_dp, vp]
        L.gsl_rng_alloc.restype = vp
        L.gsl_rng_alloc.argtypes = [vp]
        L.gsl_stub_feed.argtypes = [_dp, C.c_int, _dp, C.c_int]
        L.gsl_stub_feed_left.argtypes = [C.c_int]
        L.run_chain.argtypes = [vp, C.c_int, vp, _dp, _dp, vp, vp, C.c_int, vp, vp, vp, _dp, _dp, _dp, C.c_double, _dp,
                                _dp, _dp, vp, vp]
        L.ptmcmc.argtypes = [vp, _dp, _dp]
        self.rng = L.gsl_rng_alloc(None)
        self.limited = (C.c_double * (2 * GAIA_NPARS))()
        self.limits = (C.c_double * (2 * GAIA_NPARS))()
        self.gauss = (C.c_int * GAIA_NPARS)()
        L.set_limits(self.limited, self.limits, self.gauss)

    def set_limits(self):
        ld = np.array(self.limited).reshape(GAIA_NPARS, 2)
        lm = np.array(self.limits).reshape(GAIA_NPARS, 2)
        return lm[:, 0].copy(), lm[:, 1].copy(), ld[:, 0].copy(), ld[:, 1].copy(), np.array(self.gauss, dtype=np.int32)

    def proposal_sigmas(self):
        """init_proposals writes sigma[0..1] only; the rest is whatever the caller's buffer held."""
        s = np.full(GAIA_NPARS, np.nan)
        self.lib.init_proposals(_p(s), None)
        return s

    def gaussian(self, x, m, s): return self.lib.gaussian(x, m, s)

    def get_logP(self, pars):
        return self.lib.get_logP(_p(_f64(pars).copy()), self.limited, self.limits, self.gauss)

    def get_mags(self, p6, D):
        out = np.empty(4)
        self.lib.get_mags(_p(_f64(p6).copy()), D, _p(out))
        return out

    def model_likelihood(self, data, err, p6, D):
        model = np.empty(4)
        return self.lib.model_likelihood(_p(_f64(data).copy()), _p(_f64(err).copy()), _p(model), _p(_f64(p6).copy()), D)

    def run_chain(self, it, x, sigma, temp, index, history, chain_id, data, err, D, logLx, uniforms, normals):
        """One call of run_chain (GAIA_mcmc.c:518-590) for rung `chain_id`, with the gsl draws fed.
        x[NCHAINS][6], history[NCHAINS][NPAST][6], logLx[NCHAINS], index[NCHAINS] are updated in
        place; returns (y, accepted, DE accepted, draws left over)."""
        n, T = GAIA_NPARS, x.shape[0]
        rows = lambda a: (C.c_void_p * a.shape[0])(*[a[i].ctypes.data for i in range(a.shape[0])])
        xr = rows(x)
        y = np.zeros((T, n))
        yr = rows(y)
        hrows = [rows(history[j]) for j in range(T)]
        hr = (C.c_void_p * T)(*[C.addressof(h) for h in hrows])
        detrial = (C.c_int * T)()
        acc = np.zeros(T)
        deacc = np.zeros(T)
        model = np.empty(4)
        u = _f64(uniforms).copy()
        z = _f64(normals).copy()
        self.lib.gsl_stub_feed(_p(u), u.size, _p(z), z.size)
        self.lib.run_chain(self.rng, int(it), xr, _p(sigma), _p(temp), index.ctypes.data, hr, int(chain_id), detrial,
                           self.limits, self.limited, _p(data), _p(err), _p(model), float(D), _p(acc), _p(deacc),
                           _p(logLx), yr, self.gauss)
        left = (self.lib.gsl_stub_feed_left(0), self.lib.gsl_stub_feed_left(1))
        self.lib.gsl_stub_feed(None, 0, None, 0)
        return y[chain_id].copy(), left

    def ptmcmc(self, index, temp, logL):
        self.lib.ptmcmc(index.ctypes.data, _p(temp), _p(logL))

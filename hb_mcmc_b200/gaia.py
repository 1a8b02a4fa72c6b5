"""Stand-alone Gaia-colour sampler (host-side mirror of the hb_gaia_pt_* C ABI).

The reference program is GAIA_mcmc.c: ``./gaia NITER TIC NTHREADS`` runs 20 tempered chains of 6
parameters {logM1, logM2, rr1, rr2, aT1, aT2} against one star's {G, B-V, V-G, G-T} (run_mcmc,
GAIA_mcmc.c:663-780).  On the device a whole run is ONE kernel launch and any number of independent
ladders (``n_ens``: other stars, or other seeds of the same star) run side by side, a warp each.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .lib import Context, HBError, _f64, _p
from .pt import COUNTER_NAMES

GAIA_NPARS = 6   # MAGPARS, GAIA_mcmc.c:26
GAIA_NCHAINS = 20  # :23
GAIA_NPAST = 100   # :24
GAIA_DTEMP = 1.2   # :476
GAIA_THIN = 10     # log every 10th iteration, :760


def read_mag_file(path: str):
    """``../data/magnitudes/<TIC>.txt`` (read_mag_data, GAIA_mcmc.c:314-343): distance, then four
    ``value<TAB>error`` lines for G, B-V, V-G, G-T."""
    with open(path) as fh:
        tok = fh.read().split()
    if len(tok) < 9:
        raise ValueError(f"{path}: expected 9 numbers (distance + 4 value/error pairs), found {len(tok)}")
    v = [float(x) for x in tok[:9]]
    return v[0], np.array(v[1::2]), np.array(v[2::2])


class GaiaSampler:
    def __init__(self, ctx: Context, n_ens: int = 1, n_temps: int = GAIA_NCHAINS, seed: int = 1,
                 dtemp: float = GAIA_DTEMP, npast: int = GAIA_NPAST):
        self.ctx, self._L = ctx, ctx._L
        self.n_ens, self.n_temps, self.npast = int(n_ens), int(n_temps), int(npast)
        self.n_walkers = self.n_ens * self.n_temps
        self.seed, self.dtemp = int(seed), float(dtemp)
        h = C.c_void_p()
        rc = self._L.hb_gaia_pt_create(ctx.handle, C.byref(h), self.n_temps, self.n_ens, C.c_ulonglong(self.seed),
                                       self.dtemp, self.npast)
        if rc != 0:
            raise HBError(self._L.hb_last_error(ctx.handle).decode() or f"hb_gaia_pt_create failed ({rc})")
        self._h = h

    def _ck(self, rc):
        if rc != 0:
            raise HBError(self._L.hb_last_error(self.ctx.handle).decode() or f"libhb_b200 error {rc}")

    def close(self):
        if getattr(self, "_h", None):
            self._L.hb_gaia_pt_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    @property
    def temps(self) -> np.ndarray:
        t = np.empty(self.n_temps)
        t[0] = 1.0
        for i in range(1, self.n_temps):
            t[i] = t[i - 1] * self.dtemp  # GAIA_mcmc.c:478-484
        return t

    @property
    def iteration(self) -> int:
        return int(self._L.hb_gaia_pt_iteration(self._h))

    def set_data(self, D, data, err):
        """Distance (pc), {G, B-V, V-G, G-T} and errors; one star (broadcast) or one per ensemble."""
        D = np.ascontiguousarray(np.broadcast_to(_f64(D).reshape(-1), (self.n_ens,)), dtype=np.float64)
        data = np.ascontiguousarray(np.broadcast_to(_f64(data).reshape(-1, 4), (self.n_ens, 4)), dtype=np.float64)
        err = np.ascontiguousarray(np.broadcast_to(_f64(err).reshape(-1, 4), (self.n_ens, 4)), dtype=np.float64)
        self._ck(self._L.hb_gaia_pt_set_data(self._h, _p(D), _p(data), _p(err)))

    def set_sigma(self, sigma6):
        s = _f64(sigma6).reshape(GAIA_NPARS).copy()
        self._ck(self._L.hb_gaia_pt_set_sigma(self._h, _p(s)))

    def init_random(self):
        self._ck(self._L.hb_gaia_pt_init_random(self._h))

    def set_state(self, x):
        x = np.ascontiguousarray(_f64(x).reshape(self.n_walkers, GAIA_NPARS))
        self._ck(self._L.hb_gaia_pt_set_state(self._h, _p(x)))

    def run(self, n_iters: int, thin: int = GAIA_THIN, log: bool = True):
        """n_iters iterations in one launch.  Returns (chain[E, R, 7], logL_by_rung[E, R, T]) with one
        record per iteration ``it % thin == 0`` (what the reference appends to its chain / logL files),
        or None when ``log`` is false."""
        n_iters = int(n_iters)
        if not log:
            self._ck(self._L.hb_gaia_pt_run(self._h, n_iters, 0, None, None))
            return None
        R = int(self._L.hb_gaia_pt_records(self._h, n_iters, int(thin)))
        chain = np.empty((self.n_ens, R, GAIA_NPARS + 1))
        by_rung = np.empty((self.n_ens, R, self.n_temps))
        self._ck(self._L.hb_gaia_pt_run(self._h, n_iters, int(thin), _p(chain), _p(by_rung)))
        return chain, by_rung

    def state(self):
        x = np.empty((self.n_walkers, GAIA_NPARS))
        logL = np.empty(self.n_walkers)
        index = np.empty((self.n_ens, self.n_temps), dtype=np.int32)
        self._ck(self._L.hb_gaia_pt_get_state(self._h, _p(x), _p(logL), index.ctypes.data_as(C.POINTER(C.c_int))))
        return x, logL, index

    def proposal(self):
        """Last iteration's proposals by rung: (y[W,6], logLy[W], logPy[W], jump[W])."""
        y = np.empty((self.n_walkers, GAIA_NPARS))
        ll, lp = np.empty(self.n_walkers), np.empty(self.n_walkers)
        jump = np.empty(self.n_walkers, dtype=np.int32)
        self._ck(self._L.hb_gaia_pt_get_proposal(self._h, _p(y), _p(ll), _p(lp), jump.ctypes.data_as(C.POINTER(C.c_int))))
        return y, ll, lp, jump

    def history(self):
        h = np.empty((self.n_walkers, self.npast, GAIA_NPARS))
        self._ck(self._L.hb_gaia_pt_get_history(self._h, _p(h)))
        return h

    def map(self):
        xm = np.empty((self.n_ens, GAIA_NPARS))
        lm = np.empty(self.n_ens)
        self._ck(self._L.hb_gaia_pt_get_map(self._h, _p(xm), _p(lm)))
        return xm, lm

    def counters(self):
        out = np.zeros((self.n_ens, 8), dtype=np.uint64)
        self._ck(self._L.hb_gaia_pt_get_counters(self._h, out.ctypes.data_as(C.POINTER(C.c_ulonglong))))
        return {k: out[:, i].copy() for i, k in enumerate(COUNTER_NAMES)}

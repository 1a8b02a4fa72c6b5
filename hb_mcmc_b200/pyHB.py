"""``pyHB`` on the B200 path -- the module surface of the reference's Cython binding
(pyHB.pyx:31-130,230-294) implemented with ctypes over the C ABI (no Cython needed).

Same call names and return conventions: ``lightcurve3``, ``calc_mags``, ``calc_radii_and_Teffs``,
``getR``, ``getT``, ``envelope_Temp``, ``envelope_Radius``, ``likelihood``, ``parspace``, ``sp3``,
``test_roche_lobe``; plus the batched ``likelihood_batch`` / ``lightcurve3_batch``.

Quirk Q9: the reference binding repacks the 21 physical parameters into a
stale 22-slot layout (dummy Omega at [5], exp() pre-applied to the beaming rescale) before calling
the 21-slot ``calc_light_curve`` (pyHB.pyx:33-36 vs likelihood3.c:533-578), which shifts every
parameter from index 5 on.  This module passes the 21 parameters in the physical order of
likelihood3.c:533-578, i.e. it computes what the C model defines.  ``lightcurve3(...,
reference_q9_layout=True)`` reproduces the stale marshalling for comparisons with old outputs.
"""
from __future__ import annotations

import numpy as np

from .lib import NPARS, Context

_ctx: Context | None = None


def context(device: int = 0) -> Context:
    """Lazily created module-level device context (raises HBError without a B200)."""
    global _ctx
    if _ctx is None:
        _ctx = Context(device)
    return _ctx


def _pars21(p):
    p = np.asarray(p, dtype=np.float64)
    if p.shape[-1] == NPARS + 1:  # trailing ln_noise_resc, as calc_mags / likelihood receive it
        p = p[..., :-1]
    if p.shape[-1] != NPARS:
        raise ValueError(f"expected {NPARS} model parameters, got {p.shape[-1]}")
    return p


def lightcurve3(times, inpars, reference_q9_layout: bool = False):
    """Model flux at `times` for the 21 parameters (pyHB.pyx:31-69); blend and flux_tune applied."""
    p = _pars21(inpars)
    if reference_q9_layout:
        q = p.copy()
        q22 = np.concatenate([q[:5], [0.0], q[5:15], np.exp(q[15:17]), q[17:]])  # pyHB.pyx:36
        p = q22[:NPARS]  # the C callee reads 21 slots of the 22-slot array
    return context().calc_light_curve(np.asarray(times, dtype=np.float64), p)


def lightcurve3_batch(times, pars):
    """[n, 21] -> [n, len(times)] on the device in one call."""
    ctx = context()
    t = np.asarray(times, dtype=np.float64)
    ctx.set_data(t, np.ones_like(t), np.ones_like(t))
    return ctx.light_curves(_pars21(pars))


def calc_mags(params, Distance):
    """[G, B-V, V-G, G-T] (pyHB.pyx:71-89); `params` may carry the trailing ln_noise_resc."""
    return list(context().chain_info(_pars21(params)[None], float(Distance))[0, 4:8])


def calc_radii_and_Teffs(params):
    """R1(Rsun), R2(Rsun), Teff1(K), Teff2(K) (pyHB.pyx:91-105)."""
    return tuple(context().chain_info(_pars21(params)[None])[0, :4])


def getT(logM):
    return context().scalar(0, float(logM))


def getR(logM):
    return context().scalar(1, float(logM))


def envelope_Temp(logM):
    return context().scalar(2, float(logM))


def envelope_Radius(logM):
    return context().scalar(3, float(logM))


def likelihood(times, fluxes, errs, pars, lctype=3):
    """Gaussian log-likelihood with a noise-rescale parameter as the last entry of `pars`
    (pyHB.pyx:230-252); failures and NaN map to -1e18."""
    minlike = -1e18
    try:
        if lctype != 3:
            raise ValueError("only lctype=3 is supported (lctype=2 is retired in the reference too)")
        ln_resc = float(pars[-1])
        model = lightcurve3(times, pars[:-1])
        sig = np.asarray(errs, dtype=np.float64) * np.exp(ln_resc)
        ll = -np.sum(((np.asarray(fluxes) - model) / sig) ** 2) / 2 - len(sig) * ln_resc
    except Exception as exc:  # the reference prints the traceback and returns the floor
        print("likelihood exception:", exc)
        ll = minlike
    return ll if ll > minlike else minlike


_batch_ctx: Context | None = None


def likelihood_batch(times, fluxes, errs, pars, device: int = 0):
    """`likelihood` (pyHB.pyx:230-252) for pars[n, 22] with one device call: the n model light curves come
    from the batched kernel, the Gaussian sum is taken on the host exactly as the scalar function takes it --
    no Roche override, no 1e-5 noise clamp (loglikelihood of likelihood3.c applies both, the binding's
    likelihood neither).  A private context is used, so the data set and magnitudes of `context()` stay as
    the caller left them."""
    global _batch_ctx
    P = np.asarray(pars, dtype=np.float64).reshape(-1, NPARS + 1)
    times = np.asarray(times, dtype=np.float64)
    fluxes, errs = np.asarray(fluxes, dtype=np.float64), np.asarray(errs, dtype=np.float64)
    if _batch_ctx is None:
        _batch_ctx = Context(device)
    _batch_ctx.set_data(times, fluxes, errs)
    models = _batch_ctx.light_curves(P[:, :NPARS])
    ln = P[:, NPARS]
    sig = errs[None, :] * np.exp(ln)[:, None]
    with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
        ll = -np.sum(((fluxes[None, :] - models) / sig) ** 2, axis=1) / 2 - errs.size * ln
    return np.where(ll > -1e18, ll, -1e18)  # NaN and anything below the floor map to the floor (pyHB.pyx:250)


class parspace:
    """Named box of parameters with optional pinned entries (interface of pyHB.pyx:133-183)."""

    def __init__(self, *args):
        if len(args) % 2:
            raise ValueError("parspace: arguments are ('name1',[min,max],'name2',[min,max],...)")
        self.names = list(args[0::2])
        rng = np.array(args[1::2], dtype=np.float64).reshape(-1, 2)
        self.mins, self.maxs = rng[:, 0].copy(), rng[:, 1].copy()
        self.N = self.Nlive = len(self.names)
        self.live = np.ones(self.N, dtype=bool)
        self.pinvals = [None] * self.N
        self.idx = {n: i for i, n in enumerate(self.names)}

    def reset_range(self, name, minmax):
        i = self.idx[name]
        if not self.live[i] and not (minmax[0] <= self.pinvals[i] <= minmax[1]):
            raise ValueError("pinned value is not within range")
        self.mins[i], self.maxs[i] = minmax

    def pin(self, name, value):
        i = self.idx[name]
        if not (self.mins[i] <= value <= self.maxs[i]):
            print(f"parspace.pin: Value {name} = {value}  out of range [{self.mins[i]},{self.maxs[i]}]")
            return False
        self.Nlive -= int(self.live[i])
        self.live[i] = False
        self.pinvals[i] = value
        return True

    def get_pars(self, livevals):
        out = np.array(self.pinvals, dtype=object)
        out[self.live] = livevals
        return out

    def live_ranges(self):
        return np.column_stack([self.mins[self.live], self.maxs[self.live]])

    def live_names(self):
        return [n for n, l in zip(self.names, self.live) if l]

    def draw_live(self):
        return np.random.rand(self.Nlive) * (self.maxs - self.mins)[self.live] + self.mins[self.live]

    def out_of_bounds(self, pars):
        p = np.asarray(pars)
        return not bool(np.all((p >= self.mins) & (p <= self.maxs)))


# the 21 model parameters + noise rescale with the ranges of pyHB.pyx:205-228
sp3 = parspace(
    "logM1", [-1.5, 2.0], "logM2", [-1.5, 2.0], "logP", [-2.0, 3.0], "e", [0, 1], "inc", [0, np.pi],
    "omega0", [-np.pi, np.pi], "T0", [-1000, 1000], "alp_rad1_resc", [-1, 1], "alp_rad2_resc", [-1, 1],
    "mu_1", [0.12, 0.20], "tau_1", [0.30, 0.38], "mu_2", [0.12, 0.20], "tau_2", [0.30, 0.38],
    "alp_ref_1", [0.8, 1.2], "alp_ref_2", [0.8, 1.2], "ln_beam_resc_1", [-0.1, 0.1], "ln_beam_resc_2", [-0.1, 0.1],
    "alp_Teff_1", [-1, 1], "alp_Teff_2", [-1, 1], "blend_frac", [0.0, 1.0], "flux_tune", [0.99, 1.01],
    "ln_noise_resc", [-0.2, 0.2])


def test_roche_lobe(pars, Roche_type="L1", verbose=False):
    """max(R/R_Hill) over the two stars at periastron (pyHB.pyx:256-294); > 1 means overflow.
    `pars` carries the trailing ln_noise_resc, and pars[2] is used as the period in days exactly
    as the reference does."""
    M1, M2 = 10 ** pars[0], 10 ** pars[1]
    q = M2 / M1
    P, e = pars[2], pars[3]
    R1, R2, _, _ = calc_radii_and_Teffs(pars[:-1])
    Rsec, Rpri = (R2, R1) if q <= 1 else (R1, R2)
    a = 4.208278 * ((M1 + M2) * P ** 2) ** (1 / 3)
    if Roche_type == "L1":
        fsec = ((q + 2 / 3 + 1 / q) * 3) ** (-1 / 3)
        fpri = 1 - fsec
    elif Roche_type == "Eggleton":
        fsec = 0.49 / (0.6 + q ** (-2 / 3) * np.log(1 + q ** (1 / 3)))
        fpri = 0.49 / (0.6 + q ** (2 / 3) * np.log(1 + q ** (-1 / 3)))
    else:
        raise ValueError(f'Did not recognize Roche_type="{Roche_type}"')
    rperi = a * (1 - e)
    if verbose:
        print("Roche lobe test: Rsec, RHillsec, Rpri, RHillpri, :", Rsec, rperi * fsec, Rpri, rperi * fpri)
    return max(Rsec / (rperi * fsec), Rpri / (rperi * fpri))


test_roche_lobe.__test__ = False  # not a pytest test

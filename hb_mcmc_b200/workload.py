"""Synthetic TESS-cadence workloads of BASELINE.json / SURVEY.md section 8(d).

Pure numpy input generation -- no model arithmetic.  Anything that needs the model (the
truth light curve, the Roche-overflow rejection of prior draws) is passed in as a callable,
so the GPU legs use the GPU library and the CPU legs use the oracle on identical inputs.
"""
from __future__ import annotations

import numpy as np

NPARS = 21
PI = 3.14159265358979323846

# test_likelihoods.c:33-36 -- the one reference-supplied parameter vector (e = 0.226, P = 2.07 d)
TRUTH_A = np.array([
    0.300918167925, 0.201073240382, 0.315687018, 0.226228332961, 1.43483081695, 2.39254328279,
    1.53065288845, -2.45048204944, -0.0111151536831, 0.17004110046, 0.334086604211, 0.155971855936,
    0.339246868468, 0.94581378228, 0.824791381832, -0.0140470354976, -0.0368550857758, 0.41323817757,
    0.547024296055, 0.308828039741, 1.00029951763])
# stress truth, e = 0.95, P = 39.8 d
TRUTH_B = np.array([0.10, 0.00, 1.60, 0.95, 1.20, 0.50, 3.0, -1.0, -1.0, 0.16, 0.34, 0.16, 0.34, 1.0, 1.0,
                    0., 0., 0., 0., 0.05, 1.0])

CADENCE_DAYS = 2.0 / 1440.0  # TESS 2-minute cadence
SIGMA = 3.0e-4


def time_grid(n_points: int) -> np.ndarray:
    return np.arange(n_points, dtype=np.float64) * CADENCE_DAYS


def prior_box(lc_period: float):
    """Bounds of set_limits (likelihood3.c:986-1121): (lo, hi) arrays of 21."""
    lo = np.array([-1.5, -1.5, -2.0, 0.0, 0.0, -PI, 0.0, -5., -5., 0.12, 0.3, 0.12, 0.3, 0.5, 0.5, -0.3, -0.3,
                   -5., -5., 0., 0.99])
    hi = np.array([2.0, 2.0, 3.0, 1.0, PI, PI, lc_period, 5., 5., 0.20, 0.38, 0.20, 0.38, 1.5, 1.5, 0.3, 0.3,
                   5., 5., 1., 1.01])
    return lo, hi


def draw_chains(n: int, truth: np.ndarray, roche_fn, seed: int = 1, e_max: float = 0.95) -> np.ndarray:
    """n prior draws with the period pinned to the truth (mcmc_wrapper2.c:478), T0 in [0, P),
    e in [0, e_max], re-drawing while roche_fn(params[k, 21]) -> int[k] flags overflow
    (precedent: test_likelihoods.c:124-134)."""
    rng = np.random.default_rng(seed)
    period = 10.0 ** truth[2]
    lo, hi = prior_box(period)
    hi = hi.copy()
    hi[3] = e_max
    out = np.empty((n, NPARS))
    filled = 0
    while filled < n:
        m = max(256, 2 * (n - filled))
        P = lo + rng.random((m, NPARS)) * (hi - lo)
        P[:, 2] = truth[2]
        ok = np.asarray(roche_fn(P)) == 0
        P = P[ok][: n - filled]
        out[filled:filled + len(P)] = P
        filled += len(P)
    return out


def make_dataset(n_points: int, truth: np.ndarray, light_curve_fn, seed: int = 20240229):
    """(t, flux, err): flux = model(truth) + SIGMA N(0,1), err = SIGMA."""
    t = time_grid(n_points)
    model = np.asarray(light_curve_fn(t, truth))
    rng = np.random.default_rng(seed)
    flux = model + SIGMA * rng.standard_normal(n_points)
    err = np.full(n_points, SIGMA)
    return t, flux, err


# BASELINE.json "configs": name -> (n_chains, n_points, truth, gaia term)
CONFIGS = {
    "C1": dict(n_chains=1, n_points=20_000, truth="A", gaia=False),
    "C2": dict(n_chains=4096, n_points=20_000, truth="A", gaia=False),
    "C3": dict(n_chains=64 * 256, n_points=20_000, truth="A", gaia=False),
    "C4": dict(n_chains=8192, n_points=50_000, truth="A", gaia=True),
    "C5": dict(n_chains=16_384, n_points=200_000, truth="B", gaia=False),
}

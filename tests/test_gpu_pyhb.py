"""pyHB surface (pyHB.pyx) on the B200 path."""
import numpy as np
import pytest

from hb_mcmc_b200 import workload as wl

pytestmark = pytest.mark.gpu


def test_pyhb_functions(golden, orc):
    from hb_mcmc_b200 import pyHB
    kat = golden["kat_params"]
    lc = pyHB.lightcurve3(golden["kat_times"], list(kat))
    assert np.abs(lc - golden["kat_lc"]).max() < 2e-12
    assert np.allclose(pyHB.calc_radii_and_Teffs(kat), golden["kat_radii_teffs"], rtol=1e-13)
    assert np.abs(np.array(pyHB.calc_mags(list(kat) + [0.0], 100.0)) - golden["kat_mags_D100"]).max() < 1e-12
    assert abs(pyHB.getT(0.3) - orc.getT(0.3)) < 1e-13 and abs(pyHB.getR(-0.2) - orc.getR(-0.2)) < 1e-13
    assert pyHB.envelope_Temp(0.1) == 0.0224 and abs(pyHB.envelope_Radius(0.1) - orc.envelope_radius(0.1)) < 1e-14
    # likelihood with the noise rescale parameter, scalar and batched
    t = wl.time_grid(1000)
    flux, err = golden["n1000_flux"], np.full(1000, wl.SIGMA)
    P = np.column_stack([golden["n1000_params"][:8], np.linspace(-0.2, 0.2, 8)])
    one = [pyHB.likelihood(t, flux, err, p) for p in P]
    batch = pyHB.likelihood_batch(t, flux, err, P)
    want = []
    for p in P:
        m = orc.calc_light_curve(t, p[:21])
        want.append(-np.sum(((flux - m) / (err * np.exp(p[21]))) ** 2) / 2 - 1000 * p[21])
    assert np.allclose(one, want, rtol=1e-10) and np.allclose(batch, want, rtol=1e-10)
    assert pyHB.likelihood(t, flux, err, P[0], lctype=2) == -1e18
    # Q9: the stale 22-slot marshalling is available for comparisons and differs from the physical layout
    stale = pyHB.lightcurve3(golden["kat_times"], list(kat), reference_q9_layout=True)
    assert np.abs(stale - lc).max() > 1e-3
    sp = pyHB.sp3
    assert sp.N == 22 and sp.names[5] == "omega0" and not sp.out_of_bounds(sp.draw_live())
    assert sp.pin("logP", 0.3) and sp.Nlive == 21 and len(sp.live_names()) == 21
    assert pyHB.test_roche_lobe(list(kat) + [0.0], "Eggleton") > 0 and pyHB.test_roche_lobe(list(kat) + [0.0]) > 0
    lcs = pyHB.lightcurve3_batch(golden["kat_times"], np.stack([kat, kat]))
    assert np.abs(lcs - golden["kat_lc"]).max() < 2e-12

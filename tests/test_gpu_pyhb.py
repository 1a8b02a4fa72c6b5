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
    # the batch is `likelihood` for every row: a Roche-overflowing row and tiny error bars get neither the
    # -5e14 override nor the 1e-5 noise clamp of loglikelihood() (pyHB.pyx:230-252 applies neither)
    Pr = np.column_stack([golden["roche_params"][:4], np.zeros(4)])
    tiny = np.full(1000, 1e-7)
    one_r = [pyHB.likelihood(t, flux, tiny, p) for p in Pr]
    assert np.allclose(pyHB.likelihood_batch(t, flux, tiny, Pr), one_r, rtol=1e-10)
    # and it leaves the module-level context (data set, magnitudes) as the caller set it
    c = pyHB.context()
    c.set_data(t, flux, err)
    c.set_mags(golden["kat_mag_data"], golden["kat_mag_err"], 1, 0)
    before = c.loglikelihood(golden["n1000_params"][:4])
    pyHB.likelihood_batch(t[:500], flux[:500], tiny[:500], Pr)
    assert np.array_equal(c.loglikelihood(golden["n1000_params"][:4]), before)
    c.set_mags([1000, 1, 1, 1, 1], [1e15] * 4, 1, 0)
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


def _load_ext(sub):
    """The UNMODIFIED pyHB.pyx compiled by `make -C oracle pyhb` (oracle/_ref, built where the reference sources
    exist; the .so travels)."""
    import glob, importlib.util, os
    here = os.path.dirname(os.path.abspath(__file__))
    so = glob.glob(os.path.join(here, "..", "oracle", "_ref", sub, "pyHB*.so"))
    if not so:
        pytest.skip(f"oracle/_ref/{sub} not built (make -C oracle pyhb)")
    spec = importlib.util.spec_from_file_location("pyHB", so[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_reference_cython_binding_on_the_shim(golden):
    """pyHB.pyx itself (not our Python mirror) linked against libhb_likelihood3.so through the .pxd of
    INTEGRATION.md section 3, next to the same .pyx on the reference's likelihood3.c: same answers through the
    binding's own marshalling (22-slot Q9 layout included), so a user of `import pyHB` sees no change."""
    ref, dev = _load_ext("pyhb_ref"), _load_ext("pyhb_shim")
    kat = list(golden["kat_params"])
    t = golden["kat_times"]
    a, b = np.asarray(ref.lightcurve3(t, kat)), np.asarray(dev.lightcurve3(t, kat))
    assert a.shape == b.shape == (len(t),) and np.isfinite(a).all()
    assert np.abs(a - b).max() < 2e-12
    pars22 = kat + [0.0]
    assert np.allclose(dev.calc_radii_and_Teffs(kat), ref.calc_radii_and_Teffs(kat), rtol=1e-13)
    assert np.allclose(dev.calc_mags(pars22, 100.0), ref.calc_mags(pars22, 100.0), rtol=0, atol=1e-12)
    for m in (-1.2, -0.2, 0.0, 0.3, 1.1):
        assert abs(dev.getT(m) - ref.getT(m)) < 1e-13 and abs(dev.getR(m) - ref.getR(m)) < 1e-13
        assert dev.envelope_Temp(m) == ref.envelope_Temp(m)
        assert abs(dev.envelope_Radius(m) - ref.envelope_Radius(m)) < 1e-14
    flux, err = a + 1e-4, np.full(len(t), 3e-4)
    la, lb = ref.likelihood(t, flux, err, pars22), dev.likelihood(t, flux, err, pars22)
    assert abs(la - lb) <= 1e-10 * abs(la)

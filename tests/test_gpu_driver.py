"""host/hb_mcmc: the reference's CLI and on-disk formats on the device sampler."""
import os
import subprocess

import numpy as np
import pytest

from hb_mcmc_b200 import build
from hb_mcmc_b200 import workload as wl

pytestmark = pytest.mark.gpu


def test_driver_cli_and_file_formats(tmp_path, ctx):
    build.build_lib()
    exe = build.build_driver()
    prefix = tmp_path / "data"
    for d in ("chains", "logL", "log", "pars", "subpars", "magnitudes", "lightcurves/folded_lightcurves",
              "lightcurves/mcmc_lightcurves"):
        (prefix / d).mkdir(parents=True)
    N = 375  # size of the reference's real folded light curves (163-763 points)
    t = np.sort(np.random.default_rng(0).uniform(0, 10 ** wl.TRUTH_A[2], N))
    flux = ctx.calc_light_curve(t, wl.TRUTH_A) + 3e-4 * np.random.default_rng(1).standard_normal(N)
    with open(prefix / "lightcurves/folded_lightcurves/TIC42_new.txt", "w") as f:  # helpful_functions.py:197-203
        f.write(f"{N}\n")
        for a, b in zip(t, flux):
            f.write(f"{a:.10f}\t{b:.10f}\t{3e-4:.10f}\n")
    G = ctx.chain_info(wl.TRUTH_A[None], 100.0)[0, 4]
    with open(prefix / "magnitudes/TIC42.txt", "w") as f:  # src/README.txt:21-27
        f.write(f"100.0\n{G:.6f}\t0.05\n0.2\t0.1\n0.1\t0.1\n0.0\t0.1\n")
    env = dict(os.environ, HB_DATA_PREFIX=str(prefix), HB_NTEMPS="12", HB_SEED="3")
    r = subprocess.run([exe, "450", "TIC42", repr(float(wl.TRUTH_A[2])), "1"], env=env, capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Using color / GMAG information" in r.stdout and "PT steps/s" in r.stdout
    sfx = "TIC42_gmag_B200_1"
    chain = np.loadtxt(prefix / f"chains/chain.{sfx}.dat", ndmin=2)
    assert chain.shape == (5, 23)  # logged after iterations 0,100,...,400: "iter/10 logL p0..p20"
    assert chain[:, 0].tolist() == [0, 10, 20, 30, 40]
    assert np.all(chain[:, 4] == wl.TRUTH_A[2])  # the period is pinned (mcmc_wrapper2.c:478)
    logL = np.loadtxt(prefix / f"logL/logL.{sfx}.dat", ndmin=2)
    assert logL.shape == (5, 13) and np.allclose(logL[:, 1], chain[:, 1])
    with open(prefix / f"lightcurves/mcmc_lightcurves/{sfx}.out") as f:
        assert int(f.readline()) == N
        out = np.loadtxt(f)
    assert out.shape == (N, 3) and np.allclose(out[:, 0], t, rtol=1e-5) and np.allclose(out[:, 1], flux, rtol=1e-5)
    par = np.loadtxt(prefix / f"pars/par.{sfx}.dat")
    sub = np.loadtxt(prefix / f"subpars/subpar.{sfx}.dat")
    assert par.shape == (21,) and sub.shape == (21,)
    # the sampler improves on its random start
    assert chain[-1, 1] >= chain[0, 1]
    # missing light-curve file: the reference prints and exits 0 (mcmc_wrapper2.c:279-283)
    r = subprocess.run([exe, "10", "NOPE", "0.3", "1"], env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "Lightcurve datafile not found" in r.stdout


def test_gaia_driver_cli_and_file_formats(tmp_path, ctx):
    """host/hb_gaia_mcmc: `NITER TIC NTHREADS` and the four files of GAIA_mcmc.c (create_log_files
    :397-426, log_data :595-636) -- the layout oracle/_ref/gaia_mcmc_ref produces."""
    build.build_lib()
    build.build_driver()
    exe = build.GAIA_DRIVER
    prefix = tmp_path / "data"
    for d in ("chains", "logL", "subpars", "GAIA_runs", "magnitudes"):
        (prefix / d).mkdir(parents=True)
    with open(prefix / "magnitudes/186260283.txt", "w") as f:  # read_mag_data, :314-343
        f.write("234.296\n7.16094512\t0.0230834782584296\n-0.0066265000000005\t0.0367165032930016\n"
                "0.0212387299999997\t0.0586013200752551\n-0.0066558899999999\t0.0086725406204692\n")
    env = dict(os.environ, HB_DATA_PREFIX=str(prefix), HB_GAIA_BLOCK="700", HB_NENS="3")
    r = subprocess.run([exe, "2005", "186260283", "4"], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Opening magnitude file" in r.stdout and "initial chi2" in r.stdout and "Iter: 0 \t Best likelihood" in r.stdout
    raw = open(prefix / "chains/186260283_GAIA_run.txt").read().splitlines()
    assert len(raw) == 201 and all(ln.endswith("\t") and len(ln.split("\t")) == 8 for ln in raw)  # its 0,10,...,2000
    chain = np.loadtxt(prefix / "chains/186260283_GAIA_run.txt", ndmin=2)
    rung = np.loadtxt(prefix / "logL/186260283_GAIA_run.txt", ndmin=2)
    assert chain.shape == (201, 7) and rung.shape == (201, 20)
    assert np.allclose(chain[:, 0], rung[:, 0], rtol=1e-9)
    assert np.all(chain[:, 1:3] >= -1.5) and np.all(chain[:, 1:3] <= 2.0) and np.all(np.abs(chain[:, 3:]) <= 3.0)
    sub = np.loadtxt(prefix / "subpars/186260283_GAIA_run.txt")
    mags = np.loadtxt(prefix / "GAIA_runs/186260283_GAIA_run.txt")
    assert sub.shape == (6,) and np.allclose(sub, chain[-1, 1:], rtol=1e-9) and mags.shape == (4,)
    assert np.allclose(mags, ctx.gaia(sub[None], 234.296)[0], rtol=1e-8)
    assert chain[-40:, 0].mean() > chain[:5, 0].mean()  # the cold rung climbs from its random start
    for e in (1, 2):
        other = np.loadtxt(prefix / f"chains/186260283_GAIA_run.ens{e}.txt", ndmin=2)
        assert other.shape == (201, 7) and not np.array_equal(other, chain)
    # missing magnitude file: the reference prints and carries on with garbage; the driver stops
    r = subprocess.run([exe, "10", "NOPE", "1"], env=env, capture_output=True, text=True, timeout=120)
    assert "Could not open magnitude file" in r.stdout


def _driver_files(prefix, sfx):
    names = (f"chains/chain.{sfx}.dat", f"logL/logL.{sfx}.dat", f"lightcurves/mcmc_lightcurves/{sfx}.out", f"pars/par.{sfx}.dat",
             f"subpars/subpar.{sfx}.dat")
    return [open(prefix / n, "rb").read() for n in names]


def test_driver_on_several_devices_writes_the_same_files(tmp_path, ctx):
    """HB_DEVICES: one process, one context and host thread per device (mcmc_wrapper2.c main, :8-699, has one
    ladder and one OpenMP team).  ONE ladder over three contexts -- the rungs' likelihood evaluations split, the
    logL vector exchanged every step -- and four ladders over three contexts (whole ladders each) write the files
    of the one-device run byte for byte.  The contexts share GPU 0 here (peer-copy exchange; NCCL needs distinct
    GPUs: tests/test_gpu_multi.py and `HB_DEVICES=0,...,7 ./hb_mcmc` on a multi-GPU box)."""
    build.build_lib()
    exe = build.build_driver()
    prefix = tmp_path / "data"
    for d in ("chains", "logL", "log", "pars", "subpars", "magnitudes", "lightcurves/folded_lightcurves",
              "lightcurves/mcmc_lightcurves"):
        (prefix / d).mkdir(parents=True)
    N = 1500
    t = np.sort(np.random.default_rng(0).uniform(0, 10 ** wl.TRUTH_A[2], N))
    flux = ctx.calc_light_curve(t, wl.TRUTH_A) + 3e-4 * np.random.default_rng(1).standard_normal(N)
    with open(prefix / "lightcurves/folded_lightcurves/TIC7_new.txt", "w") as f:
        f.write(f"{N}\n")
        for a, b in zip(t, flux):
            f.write(f"{a:.10f}\t{b:.10f}\t{3e-4:.10f}\n")
    sfx = "TIC7_gmag_B200_2"
    logp = repr(float(wl.TRUTH_A[2]))

    def run(**env):
        e = dict(os.environ, HB_DATA_PREFIX=str(prefix), HB_SEED="5", **env)
        r = subprocess.run([exe, "230", "TIC7", logp, "2"], env=e, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout + r.stderr
        return r.stdout, _driver_files(prefix, sfx), open(prefix / f"log/log.{sfx}.dat").read()

    _, one, _ = run(HB_NTEMPS="10")
    out, three, log = run(HB_NTEMPS="10", HB_DEVICES="0,0,0", HB_EXCHANGE="peer")
    assert "3 devices: rungs split" in out and "split rungs" in log
    assert three == one
    _, one4, _ = run(HB_NTEMPS="6", HB_NENS="4")
    out, three4, log = run(HB_NTEMPS="6", HB_NENS="4", HB_DEVICES="0,0,0")
    assert "whole ladders per device" in out and "MAP logL by ensemble:" in log
    assert three4 == one4

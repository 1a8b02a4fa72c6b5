"""Summarise runs of the UNMODIFIED reference Gaia-colour sampler (oracle/_ref/gaia_mcmc_ref, built by
`make -C oracle ref`: GAIA_mcmc.c compiled where it lies, linked against oracle/gsl_stub because GSL
is absent here -- it supplies random numbers only) into tests/golden/gaia_reference_runs.json.

The star is TIC 186260283 of the reference's data/color_mag/cp_data_4-21-2022.csv (columns dist,
Gmag0, BmV0, VmG0, GmT0 and their errors), written in the ../data/magnitudes/<TIC>.txt format that
read_mag_data expects (GAIA_mcmc.c:314-343).  The reference seeds its generators with NITER (:676), so
neighbouring iteration counts are independent runs.

The runs use NTHREADS = 1.  With more threads the reference has a data race: every OpenMP thread
computes its magnitudes into the SAME `model` buffer (run_mcmc allocates one, :669,709, and hands it
to every run_chain call, :733-737; model_likelihood writes then reads it, :258-266), so a rung's chi^2
is now and then taken against another rung's model.  Measured here with 4 threads: the hot rungs'
mean logL drifts from -43.9 (1 thread) to anywhere between -65 and -212; the cold rung moves by a few
tenths.  The single-thread run is the algorithm as written, and the one the device sampler matches.

    make -C oracle ref && python tests/golden/make_gaia_golden.py
"""
import json
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
SCR = os.path.join(ROOT, "oracle", "_ref", "gaia_scratch")
BIN = os.path.join(ROOT, "oracle", "_ref", "gaia_mcmc_ref")
TIC = "186260283"
MAG = "234.296\n7.16094512\t0.0230834782584296\n-0.0066265000000005\t0.0367165032930016\n" \
      "0.0212387299999997\t0.0586013200752551\n-0.0066558899999999\t0.0086725406204692\n"
NITER0, NRUNS, BURN = 400000, 8, 0.25


def summarise(chain, rung):
    """Label-symmetric summaries (the likelihood is symmetric under exchanging the two stars)."""
    n0 = int(len(chain) * BURN)
    c, r = chain[n0:], rung[n0:]
    big = c[:, 1:3].argmax(axis=1)
    rows = np.arange(len(c))
    m_hi, m_lo = c[rows, 1 + big], c[rows, 2 - big]
    rr_hi, rr_lo = c[rows, 3 + big], c[rows, 4 - big]
    at_hi, at_lo = c[rows, 5 + big], c[rows, 6 - big]
    q = lambda v: np.quantile(v, [0.16, 0.5, 0.84]).tolist()
    return {"cold_logL_mean": float(c[:, 0].mean()), "cold_logL_q": q(c[:, 0]), "rung_logL_mean": r.mean(axis=0).tolist(),
            "m_hi_q": q(m_hi), "m_lo_q": q(m_lo), "rr_hi_q": q(rr_hi), "rr_lo_q": q(rr_lo), "at_hi_q": q(at_hi),
            "at_lo_q": q(at_lo), "n": int(len(c))}


def main():
    os.makedirs(os.path.join(SCR, "src"), exist_ok=True)
    with open(os.path.join(SCR, "data", "magnitudes", f"{TIC}.txt"), "w") as f:
        f.write(MAG)
    runs = []
    for k in range(NRUNS):
        niter = NITER0 + k
        subprocess.run([BIN, str(niter), TIC, "1"], cwd=os.path.join(SCR, "src"), stdout=subprocess.DEVNULL, check=False)
        chain = np.loadtxt(os.path.join(SCR, "data", "chains", f"{TIC}_GAIA_run.txt"), ndmin=2)
        rung = np.loadtxt(os.path.join(SCR, "data", "logL", f"{TIC}_GAIA_run.txt"), ndmin=2)
        s = summarise(chain, rung)
        s["niter"] = niter
        runs.append(s)
        print(niter, round(s["cold_logL_mean"], 3), [round(v, 3) for v in s["m_hi_q"]], [round(v, 3) for v in s["m_lo_q"]],
              [round(v, 2) for v in s["rr_hi_q"]], [round(v, 2) for v in s["at_hi_q"]])
    out = {"tic": TIC, "distance": 234.296, "nchains": 20, "npast": 100, "thin": 10, "burn": BURN,
           "note": "unmodified GAIA_mcmc.c, gcc -O3 -std=c99 -fopenmp, 1 thread (race-free), gsl stub RNG (splitmix64)", "runs": runs}
    with open(os.path.join(ROOT, "tests", "golden", "gaia_reference_runs.json"), "w") as f:
        json.dump(out, f)


if __name__ == "__main__":
    main()

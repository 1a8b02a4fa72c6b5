"""Summarise runs of the UNMODIFIED reference driver (oracle/_ref/hb_mcmc_ref, built by
`make -C oracle ref_driver`: mcmc_wrapper2.c with only its /scratch prefix and thread count
sed-patched) on the real folded light curve TIC 102289966 (375 points, P = 6.252443 d,
data/lightcurves/periods.txt) into tests/golden/pt_reference_runs.json.

    make -C oracle ref_driver
    cp /root/reference/data/lightcurves/folded_lightcurves/102289966_new.txt oracle/_ref/scratch/data/lightcurves/folded_lightcurves/
    cd oracle/_ref && for r in 1 2 3 4; do ./hb_mcmc_ref 20000 102289966 0.7960497 $r > scratch/run_20000_$r.log; done
    python tests/golden/make_pt_golden.py

The light curve itself (data, not code) is stored next to the summary so that the GPU test can
sample the same posterior.
"""
import json
import os
import re
import shutil

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
SCR = os.path.join(ROOT, "oracle", "_ref", "scratch")
TIC, LOGP, NITER = "102289966", 0.7960497, 20000


def main():
    runs = []
    for r in (1, 2, 3, 4):
        chain = os.path.join(SCR, "data", "chains", f"chain.{TIC}_gmag_OMP_{r}.dat")
        log = os.path.join(SCR, f"run_{NITER}_{r}.log")
        if not (os.path.exists(chain) and os.path.exists(log)):
            continue
        c = np.loadtxt(chain, ndmin=2)
        acc, deacc = [], []
        for m in re.finditer(r"(\d+)/\d+ logL=(\S+) acc=(\S+) DEacc=(\S+)", open(log).read()):
            if int(m.group(1)) >= 1000:
                acc.append(float(m.group(3)))
                if int(m.group(1)) > 1000:
                    deacc.append(float(m.group(4)))
        runs.append({"run": r, "iter": (c[:, 0] * 10).astype(int).tolist(), "cold_logL": c[:, 1].tolist(),
                     "acc_prints": acc, "deacc_prints": deacc, "final_pars": c[-1, 2:].tolist()})
    out = {"tic": TIC, "log10_period": LOGP, "niter": NITER, "nchains": 50, "npast": 500,
           "note": "unmodified reference driver, gcc -O3 -std=c99 -fopenmp, 8 threads; acc/DEacc are the driver's own prints",
           "runs": runs}
    with open(os.path.join(ROOT, "tests", "golden", "pt_reference_runs.json"), "w") as f:
        json.dump(out, f)
    shutil.copy(os.path.join(SCR, "data", "lightcurves", "folded_lightcurves", f"{TIC}_new.txt"),
                os.path.join(ROOT, "tests", "golden", f"lc_{TIC}_new.txt"))
    for r in runs:
        ll = np.array(r["cold_logL"])
        print(r["run"], len(ll), "logL@1000/5000/10000/19900:", [round(ll[i], 1) for i in (10, 50, 100, 199)],
              "acc", np.mean(r["acc_prints"]).round(3), "DEacc", np.mean(r["deacc_prints"]).round(4))


if __name__ == "__main__":
    main()

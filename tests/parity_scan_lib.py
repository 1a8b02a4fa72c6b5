"""Seeded wide parity scan of the CUDA likelihood against the compiled reference (test infrastructure).

One pass = five chain sets (size / truth / eccentricity range), half of each set within 1e-3 of the truth,
where chi^2 ~ N is most sensitive.  Shared by tests/test_gpu_parity_scan.py (the stated-bound test),
tests/tools/parity_scan.py (the campaign tool that wrote profiles/r1_parity_scan_final_kernel.txt) and
tests/golden/make_outliers.py (which pins the chains above the 1e-10 gate as fixtures).
"""
from __future__ import annotations

import numpy as np

from hb_mcmc_b200 import workload as wl

# (truth name, N, e_max, chains, base seed)
SETS = (("A", 20000, 0.95, 512, 1), ("B", 20000, 0.99, 512, 2), ("A", 1001, 0.99, 1024, 3),
        ("B", 50000, 0.97, 128, 4), ("A", 375, 0.9, 2048, 5))
TRUTHS = {"A": wl.TRUTH_A, "B": wl.TRUTH_B}


def chain_set(truth_name: str, N: int, emax: float, n: int, seed: int, roche_fn):
    """The chain set of one (set, seed): prior draws + near-truth perturbations, Roche draws removed.
    `roche_fn(P[k, 21]) -> int[k]` (the flags are exact on every implementation, so any of them will do)."""
    truth = TRUTHS[truth_name]
    P = wl.draw_chains(n, truth, roche_fn, seed=seed, e_max=emax)
    P[0] = truth
    k = n // 2
    P[1:k] = truth + 1e-3 * np.random.default_rng(seed).standard_normal((k - 1, 21)) * np.abs(truth + 0.1)
    P[1:k, 2] = truth[2]
    return P[np.asarray(roche_fn(P)) == 0]


def scan(ctx, checker, reps, seed_offset: int = 0, sets=SETS, report=None):
    """Run the passes `reps` (iterable of pass numbers).  Returns (rel[all chains], records) where records
    lists every chain above 5e-11 with everything needed to evaluate it again."""
    rels, records = [], []
    for rep in reps:
        for truth_name, N, emax, n, seed0 in sets:
            seed = seed0 + 100 * rep + seed_offset
            truth = TRUTHS[truth_name]
            t, fl, er = wl.make_dataset(N, truth, checker.calc_light_curve)
            ctx.set_data(t, fl, er)
            P = chain_set(truth_name, N, emax, n, seed, ctx.roche_overflow)
            g = ctx.loglikelihood(P)
            o = checker.loglikelihood_batch(t, fl, er, P)
            if not np.array_equal(np.isnan(g), np.isnan(o)):
                raise AssertionError(f"NaN pattern differs (set {truth_name} N={N} seed={seed})")
            with np.errstate(invalid="ignore", divide="ignore"):
                rel = np.abs(g - o) / np.abs(o)
            rel = np.where(np.isnan(o), 0.0, rel)
            rels.append(rel)
            for i in np.nonzero(rel > 5e-11)[0]:
                records.append({"truth": truth_name, "N": int(N), "seed": int(seed), "chain": int(i), "rel": float(rel[i]),
                                "logL_ref": float(o[i]).hex(), "logL_gpu": float(g[i]).hex(),
                                "params": [float(v).hex() for v in P[i]]})
            if report is not None:
                i = int(np.argmax(rel))
                report(f"truth {truth_name} N={N:6d} n={len(P):5d} emax={emax}: max rel {rel.max():.3e} (e={P[i,3]:.3f}, "
                       f"logL={o[i]:.4g})  median {np.median(rel):.2e}  nan {int(np.isnan(o).sum())}")
    return np.concatenate(rels), records

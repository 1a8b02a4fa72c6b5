"""Writes tests/golden/contact_chain_v1.json: the chain of the seeded parity scan (set B, N = 20 000, seed 5902,
chain 309) whose sample 18514 lies 1.2e-8 dc from the contact d = dc = sqrt(R1^2 - R2^2) of a 33 Rsun / 0.10 Rsun
pair, where the rounding noise of h^2 decides whether asin(h/R2) is NaN (quirk Q10).  The reference value is the
compiled reference's (oracle/_ref); run in the build container:  python tests/golden/make_contact_chain.py"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
import oracle  # noqa: E402
import parity_scan_lib as ps  # noqa: E402
from hb_mcmc_b200 import workload as wl  # noqa: E402

R = oracle.Reference()
truth_name, N, emax, n, seed, chain = "B", 20000, 0.99, 512, 5902, 309
t, fl, er = wl.make_dataset(N, ps.TRUTHS[truth_name], R.calc_light_curve)
P = ps.chain_set(truth_name, N, emax, n, seed, lambda Q: np.array([R.roche_overflow(q) for q in np.atleast_2d(Q)]))
p = P[chain]
ll = R.loglikelihood_batch(t, fl, er, p[None])[0]
lc = R.calc_light_curve(t, p)
assert np.isfinite(ll) and np.isfinite(lc).all()
rec = {"truth": truth_name, "N": N, "seed": seed, "chain": chain, "sample": 18514, "logL_ref": float(ll).hex(),
       "lc_ref_at_sample": float(lc[18514]).hex(), "params": [float(v).hex() for v in p]}
json.dump([rec], open(os.path.join(HERE, "contact_chain_v1.json"), "w"), indent=1)
print(ll, lc[18514], p[3])

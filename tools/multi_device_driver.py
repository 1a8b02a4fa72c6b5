"""host/hb_mcmc on several GPUs of one box (needs >= 2 GPUs): ONE ladder with the rungs' likelihood evaluations split
over the devices and the logL vector all-gathered by NCCL from inside the captured step, against the same run on one
device -- same files byte for byte, and the steps/s of both.
    python tools/multi_device_driver.py [n_devices] [n_points] [n_temps] [n_iter]"""
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, ".")
import hb_mcmc_b200 as hb  # noqa: E402
from hb_mcmc_b200 import build, workload as wl  # noqa: E402

ndev = int(sys.argv[1]) if len(sys.argv) > 1 else 2
N = int(sys.argv[2]) if len(sys.argv) > 2 else 200000
n_temps = int(sys.argv[3]) if len(sys.argv) > 3 else 64
n_iter = int(sys.argv[4]) if len(sys.argv) > 4 else 301
truth = wl.TRUTH_B if N > 50000 else wl.TRUTH_A
exe = build.build_driver()
ctx = hb.Context(0)
prefix = tempfile.mkdtemp(prefix="hb_multi_")
for d in ("chains", "logL", "log", "pars", "subpars", "magnitudes", "lightcurves/folded_lightcurves", "lightcurves/mcmc_lightcurves"):
    os.makedirs(os.path.join(prefix, d))
t, flux, err = wl.make_dataset(N, truth, ctx.calc_light_curve)
with open(os.path.join(prefix, "lightcurves/folded_lightcurves/TIC9_new.txt"), "w") as f:
    f.write(f"{N}\n")
    for a, b, c in zip(t, flux, err):
        f.write(f"{a:.10f}\t{b:.10f}\t{c:.10f}\n")
ctx.close()
sfx = "TIC9_gmag_B200_1"
names = (f"chains/chain.{sfx}.dat", f"logL/logL.{sfx}.dat", f"lightcurves/mcmc_lightcurves/{sfx}.out", f"pars/par.{sfx}.dat")


def run(**env):
    e = dict(os.environ, HB_DATA_PREFIX=prefix, HB_SEED="5", HB_NTEMPS=str(n_temps), **env)
    t0 = time.perf_counter()
    r = subprocess.run([exe, str(n_iter), "TIC9", repr(float(truth[2])), "1"], env=e, capture_output=True, text=True, timeout=1500)
    dt = time.perf_counter() - t0
    if r.returncode != 0:
        print(r.stdout[-2000:], r.stderr[-2000:])
        raise SystemExit(1)
    rate = [l for l in r.stdout.splitlines() if l.startswith("done:")][-1]
    return [open(os.path.join(prefix, n), "rb").read() for n in names], rate, dt


one, rate1, _ = run()
print("1 device :", rate1, flush=True)
for k in sorted({2, ndev}):
    if k > ndev:
        continue
    files, rate, _ = run(HB_DEVICES=",".join(str(i) for i in range(k)))
    print(f"{k} devices:", rate, " files identical to the one-device run:", files == one, flush=True)
    assert files == one
if ndev >= 2:
    files, rate, _ = run(HB_DEVICES=",".join(str(i) for i in range(ndev)), HB_EXCHANGE="peer")
    print(f"{ndev} devices (peer copies):", rate, " identical:", files == one, flush=True)

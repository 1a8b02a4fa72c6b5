"""SASS of the sample loop of k_chain_eval from an `ncu --set full --import-source on` capture, in address order, every
instruction with the source line it belongs to and its executions per warp-point (1.0 = once per sample):
    python tools/ncu_hotloop.py gpurun_out/prof_r2_chain_eval.ncu-rep 81920000 > profiles/r2_chain_eval_hot_loop.sass"""
import csv
import os
import subprocess
import sys

rep = sys.argv[1]
npts = float(sys.argv[2]) if len(sys.argv) > 2 else 4096 * 20000
pts = npts / 32


def page(view):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", view], capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))


# address -> file:line from the correlated view (one block per source file)
where = {}
fpath, line = "", ""
for r in page("cuda,sass"):
    if len(r) == 2 and r[0] == "File Path":
        fpath = os.path.basename(r[1])
    elif len(r) > 4 and r[0].isdigit():
        line = r[0]  # a source line; the SASS rows correlated with it follow
    elif len(r) > 4 and r[0] == "" and r[2].startswith("0x"):
        where.setdefault(r[2], f"{fpath}:{line}")
rows = page("sass")
hdr, data = rows[1], rows[2:]
iA, iS, iE = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed")
ex = []
for r in data:
    try:
        ex.append(int(r[iE]) / pts)
    except Exception:
        ex.append(0.0)
# the loop: from the first to the last instruction executed about once per sample, around the 16-byte data load
ldg = [i for i, r in enumerate(data) if "LDG.E.128" in r[iS] and ex[i] > 0.9]
if not ldg:
    sys.exit("no hot 16-byte load found")
lo = hi = ldg[0]
while lo > 0 and any(e > 0.9 for e in ex[max(0, lo - 250):lo]):
    lo -= 1
while hi + 1 < len(data) and any(e > 0.9 for e in ex[hi + 1:hi + 251]):
    hi += 1
while ex[lo] < 0.9:
    lo += 1
while ex[hi] < 0.9:
    hi -= 1
tot = 0.0
for i in range(lo, hi + 1):
    r = data[i]
    tot += ex[i]
    print(f"{i:6d} {ex[i]:6.3f}  {r[iS].strip():78s} {where.get(r[iA], '')}")
print(f"# {hi - lo + 1} instructions in the span, {tot:.1f} executed per warp-point")

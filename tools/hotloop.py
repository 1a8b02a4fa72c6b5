"""Static view of the model-pass loops of k_chain_eval in a compiled library or cubin (no GPU needed): every loop
(backward branch) that holds the 16-byte {flux, 1/sigma} load is listed with its size and opcode histogram, and
written in full to <out>.loopK.sass.  Cold blocks inside the span (eclipse, append, fmod repair, further Newton
steps) are part of the count -- compare builds with each other, not with the executed counts of an ncu capture.
    python tools/hotloop.py hb_mcmc_b200/csrc/libhb_b200.so [/tmp/out [shared]]"""
import collections
import re
import subprocess
import sys

path = sys.argv[1]
out = sys.argv[2] if len(sys.argv) > 2 else "/tmp/hotloop"
# the instantiation: batches that fill the grid (default), or "shared" for the one whose chains are shared by CTAs
WHICH = "k_chain_evalILi256ELb1" if len(sys.argv) > 3 and sys.argv[3] == "shared" else "k_chain_evalILi256ELb0"
sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
fn = None
ins = []  # (addr, text) of k_chain_eval<256>
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        fn = m.group(1)
        continue
    if fn and WHICH in fn:
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
addr_index = {a: i for i, (a, _) in enumerate(ins)}
loops = []
for i, (a, t) in enumerate(ins):
    m = re.search(r"\bBRA\b.*?0x([0-9a-f]+)", t)
    if m:
        tgt = int(m.group(1), 16)
        if tgt <= a and tgt in addr_index:
            loops.append((addr_index[tgt], i))
print(f"k_chain_eval<256>: {len(ins)} instructions")
k = 0
for s, e in loops:
    body = ins[s:e + 1]
    if not any("LDG.E.128" in t for _, t in body):
        continue
    # innermost only
    if any(s2 >= s and e2 <= e and (s2, e2) != (s, e) and any("LDG.E.128" in t for _, t in ins[s2:e2 + 1]) for s2, e2 in loops):
        continue
    h = collections.Counter()
    for _, t in body:
        op = t.split()[1] if t.startswith("@") else t.split()[0]
        h[op.split(".")[0]] += 1
    fp64 = sum(h[o] for o in ("DFMA", "DMUL", "DADD", "DSETP"))
    print(f"loop {k}: [{s}-{e}] {len(body)} instructions, FP64 {fp64}, other {len(body) - fp64}")
    print("   ", " ".join(f"{o}:{n}" for o, n in h.most_common(24)))
    with open(f"{out}.loop{k}.sass", "w") as f:
        for j, (a, t) in enumerate(body):
            f.write(f"{s + j:6d} {t}\n")
    k += 1

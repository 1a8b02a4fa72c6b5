import csv, subprocess, sys
rep = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/prof_chain_eval.ncu-rep"
npts = float(sys.argv[2]) if len(sys.argv) > 2 else 4096*20000
dump = len(sys.argv) > 3
out = subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","sass"],capture_output=True,text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]; data = rows[2:]
iS = hdr.index("Source"); iE = hdr.index("Instructions Executed"); iSamp = hdr.index("# Samples")
pts = npts/32
regions = []; cur = None
for i, r in enumerate(data):
    try: n = int(r[iE])
    except Exception: continue
    ratio = n/pts
    toks = r[iS].split(); op = (toks[1] if toks[0].startswith('@') else toks[0]).split('.')[0]
    if cur and abs(cur['ratio']-ratio) < 0.02*max(cur['ratio'],0.01)+0.002:
        cur['n'] += 1; cur['tot'] += ratio; cur['end'] = i; cur['samp'] += int(r[iSamp] or 0)
    else:
        cur = {'start': i, 'end': i, 'ratio': ratio, 'n': 1, 'tot': ratio, 'samp': int(r[iSamp] or 0), 'ops': {}}
        regions.append(cur)
    cur['ops'][op] = cur['ops'].get(op,0)+1
for r in regions:
    if r['tot'] > 2:
        ops = sorted(r['ops'].items(), key=lambda x:-x[1])[:9]
        print(f"[{r['start']:5d}-{r['end']:5d}] x{r['ratio']:6.2f} n={r['n']:4d} tot/pt {r['tot']:7.1f} samp {r['samp']:6d} {ops}")
if dump:
    a, b = map(int, sys.argv[3].split('-'))
    for i in range(a, b+1):
        r = data[i]
        print(i, r[iE], r[iSamp], r[iS])

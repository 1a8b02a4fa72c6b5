"""Finer view of an ncu capture of k_chain_eval: every straight-line SASS region with its executions per model
point, cost per point and stall samples (threshold on cost per point as argv[3], default 0.3)."""
import csv, subprocess, sys
rep = sys.argv[1]; npts = float(sys.argv[2]) if len(sys.argv) > 2 else 4096*20000
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.3
out = subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","sass"],capture_output=True,text=True).stdout
rows = list(csv.reader(out.splitlines())); hdr=rows[1]; data=rows[2:]
iS=hdr.index("Source"); iE=hdr.index("Instructions Executed"); iSamp=hdr.index("# Samples")
pts=npts/32; regions=[]; cur=None
for i,r in enumerate(data):
    try: n=int(r[iE])
    except Exception: continue
    ratio=n/pts; toks=r[iS].split(); op=(toks[1] if toks[0].startswith('@') else toks[0]).split('.')[0]
    if cur and abs(cur['ratio']-ratio) < 0.02*max(cur['ratio'],0.01)+0.002:
        cur['n']+=1; cur['tot']+=ratio; cur['end']=i; cur['samp']+=int(r[iSamp] or 0)
    else:
        cur={'start':i,'end':i,'ratio':ratio,'n':1,'tot':ratio,'samp':int(r[iSamp] or 0),'ops':{}}; regions.append(cur)
    cur['ops'][op]=cur['ops'].get(op,0)+1
tot=sum(r['tot'] for r in regions); samp=sum(r['samp'] for r in regions)
for r in regions:
    if r['tot']>thr or r['samp'] > 0.004*samp:
        ops=sorted(r['ops'].items(),key=lambda x:-x[1])[:6]
        print(f"[{r['start']:5d}-{r['end']:5d}] x{r['ratio']:7.4f} n={r['n']:4d} tot/pt {r['tot']:6.1f} samp {r['samp']:6d} {ops}")
print("total/pt", round(tot,1), "samples", samp)
if len(sys.argv) > 4:
    a,b=map(int,sys.argv[4].split('-'))
    for i in range(a,b+1): print(i, data[i][iE], data[i][iSamp], data[i][iS])

import csv, collections, subprocess, sys
rep = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/prof_chain_eval.ncu-rep"
npts = float(sys.argv[2]) if len(sys.argv) > 2 else 4096*20000
out = subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","sass"],capture_output=True,text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]; data = rows[2:]
iS = hdr.index("Source"); iE = hdr.index("Instructions Executed"); iSamp = hdr.index("# Samples")
pts = npts/32
by_op = collections.Counter(); samp = collections.Counter(); tot=0
for r in data:
    try: n = int(r[iE])
    except Exception: continue
    src = r[iS].strip(); toks = src.split()
    op = toks[1] if toks[0].startswith('@') else toks[0]
    op = op.split('.')[0]
    by_op[op]+=n; tot+=n; samp[op]+=int(r[iSamp] or 0)
print("total warp-instr per point:", round(tot/pts,1))
fp64 = sum(by_op[o] for o in ("DFMA","DMUL","DADD","DSETP"))
print("FP64 per point:", round(fp64/pts,1))
for op,n in by_op.most_common(28): print(f"{op:10s} {n/pts:7.1f}  samples {samp[op]}")
raw = subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
rows = list(csv.reader(raw.splitlines())); h,u,v = rows[0],rows[1],rows[2]
want = ["gpu__time_duration.sum","sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active","smsp__issue_active.avg.pct_of_peak_sustained_active","launch__registers_per_thread","sm__warps_active.avg.pct_of_peak_sustained_active","dram__bytes_read.sum","dram__bytes_write.sum","smsp__average_warps_issue_stalled_wait_per_issue_active.ratio","smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio","smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio","smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio","smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio","smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio","smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio","smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio","smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio","sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"]
for a,b,c in zip(h,u,v):
    if a in want: print(a,b,c)

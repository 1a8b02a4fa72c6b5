# round-2 profiling pass on one B200 (every step only after its command has exited 0 without ncu):
#   1. launch list of a short bench run (ncu --metrics gpu__time_duration.sum)  -> gpurun_out/r2_launches.csv
#   2. full capture of one k_chain_eval launch on C2                              -> gpurun_out/prof_r2_chain_eval.ncu-rep
#   3. launch list of the small-batch cases (C1, 50 x 20 000, 8 x 200 000)       -> gpurun_out/r2_launches_small.csv
set -x
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --pt-steps 8 --no-pt-reference --no-extra"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_chain_eval -s 4 -c 1 -o gpurun_out/prof_r2_chain_eval -f $CMD > gpurun_out/ncu_full.log 2>&1
tail -1 gpurun_out/ncu_full.log
SMALL="python tools/small_batches.py"
$SMALL > gpurun_out/small.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches_small.csv $SMALL > gpurun_out/ncu_small.log 2>&1
cat gpurun_out/small.log

set -x
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --pt-steps 3"
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_chain_eval -s 4 -c 1 -o gpurun_out/prof_chain_eval -f $CMD > gpurun_out/ncu_full.log 2>&1
tail -1 gpurun_out/ncu_full.log

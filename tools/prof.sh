set -x
python bench.py --steps 100 --warmup 5 > gpurun_out/bench_r1_n1.json 2> gpurun_out/bench_r1_n1.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_r1_ref.json 2>> gpurun_out/bench_r1_n1.err
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --pt-steps 3"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_chain_eval -s 4 -c 1 -o gpurun_out/prof_chain_eval -f $CMD > gpurun_out/ncu_full.log 2>&1
tail -1 gpurun_out/ncu_full.log
python tests/tools/configs.py > gpurun_out/configs.txt 2>&1; cat gpurun_out/configs.txt

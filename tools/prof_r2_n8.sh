# round-2 closing runs on eight GPUs of one box: the bench line at N = 8 and BASELINE's stress configuration C5 in full
export PYTHONPATH=.
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 100 --warmup 5 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; echo rc=$?
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --workload C5 --steps 30 --warmup 3 --no-cpu-baseline --pt-steps 0 > gpurun_out/r2_bench_c5_n8.json 2> gpurun_out/r2_bench_c5_n8.err; echo rc=$?
tail -c 400 gpurun_out/r2_bench_c5_n8.json

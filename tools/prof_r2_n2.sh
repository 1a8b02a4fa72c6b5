export PYTHONPATH=.
python -m pytest tests -m gpu -q > gpurun_out/r2_gpu_tests_n2.log 2>&1; tail -2 gpurun_out/r2_gpu_tests_n2.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo rc=$?; tail -c 300 gpurun_out/r2_bench_n2.json

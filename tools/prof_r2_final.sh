# closing run of round 2 on one B200: GPU tests, bench line (N=1), reference arm, launch lists + one full ncu capture
# (tools/prof_r2.sh), per-chain phase cycles (-DHB_PHASE_PROF variant), all configs with parity (tests/tools/configs.py)
export PYTHONPATH=.
python -m pytest tests -m gpu -q > gpurun_out/r2_gpu_tests_n1.log 2>&1; tail -2 gpurun_out/r2_gpu_tests_n1.log
python bench.py --steps 100 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo bench rc=$?
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; echo ref rc=$?
bash tools/prof_r2.sh > gpurun_out/prof_r2.log 2>&1; tail -3 gpurun_out/prof_r2.log
python tools/phase_time.py tools/variants/lib_prof.so > gpurun_out/r2_phase_cycles.txt 2>&1; cat gpurun_out/r2_phase_cycles.txt
python tests/tools/configs.py > gpurun_out/r2_configs.txt 2>&1; cat gpurun_out/r2_configs.txt

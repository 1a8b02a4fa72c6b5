#!/bin/bash
# the round's closing run on a B200 box: GPU tests, smoke, one bench line (N=1) into gpurun_out/
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 100 --warmup 5 > gpurun_out/bench_final_n1.json 2> gpurun_out/bench_final_n1.err; echo rc=$?
python - <<'PY'
import json
l = json.loads(open("gpurun_out/bench_final_n1.json").read().strip().splitlines()[-1])
print(l["value"], l["e2e"]["value"], l["ms_per_step"], l["roofline"]["kernel_ms"], l["clocks"])
print(l["pt"]["steps_per_sec"], l["pt"]["reference_size"])
print(l["cpu_baseline"])
PY

#!/bin/bash
# round 2, call B: parity suite (foreign symbols, regrowth, rounds) + c5 bench (host side of the 2-set call)
TAG=r02b
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; tail -5 gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py --workload c5 --steps 3 --warmup 1 --cpu-queries 64 > gpurun_out/${TAG}_bench_c5_1gpu.json 2> gpurun_out/${TAG}_bench_c5_1gpu.err; tail -c 600 gpurun_out/${TAG}_bench_c5_1gpu.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r02b_bench_c5_1gpu.json"))
    print({k: d[k] for k in ("value", "ms_per_step", "wall_ms_per_step", "main_kernel_ms")}, d["e2e"], d["e2e_resident"], d["roofline"]["frac"], d["cpu_baseline"])
except Exception as e:
    print("bench line unreadable:", e)
PY

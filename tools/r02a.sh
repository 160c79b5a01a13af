#!/bin/bash
# round 2, call A: parity suite + c2 bench + host-side profile of the end-to-end call
TAG=r02a
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_gpus.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; tail -5 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench_c2_1gpu.json 2> gpurun_out/${TAG}_bench_c2_1gpu.err; tail -c 600 gpurun_out/${TAG}_bench_c2_1gpu.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r02a_bench_c2_1gpu.json"))
    print({k: d[k] for k in ("value", "ms_per_step", "wall_ms_per_step", "main_kernel_ms")}, d["e2e"], d["e2e_resident"], d["roofline"]["frac"], d["cpu_baseline"])
except Exception as e:
    print("bench line unreadable:", e)
PY
timeout 300 python tools/e2e_profile.py c2 1.0 > gpurun_out/${TAG}_e2e_profile_c2.txt 2>&1; head -30 gpurun_out/${TAG}_e2e_profile_c2.txt
ISOCON_NN_DEBUG=2 timeout 300 python tools/phase_times.py c2 1.0 > gpurun_out/${TAG}_phase_times_c2.txt 2>&1; tail -40 gpurun_out/${TAG}_phase_times_c2.txt

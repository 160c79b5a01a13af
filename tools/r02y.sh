#!/bin/bash
# round 2, call Y (1 GPU): level 2 of the two-level pass as a swapped launch too (ISOCON_NN_SWAP=2) -- parity, c5 bench line
TAG=r02y
mkdir -p gpurun_out
ISOCON_NN_SWAP=2 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "two_level" > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log
ISOCON_NN_SWAP=2 ISOCON_NN_DEBUG=2 timeout 300 python tools/phase_times.py c5 1.0 > gpurun_out/${TAG}_phase_times_c5.txt 2>&1; grep -E "^rep|level 2 swapped" gpurun_out/${TAG}_phase_times_c5.txt | tail -4
ISOCON_NN_SWAP=2 timeout 600 python bench.py --workload c5 --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/${TAG}_b5.err | grep '^{' > gpurun_out/${TAG}_bench_c5_1gpu_swap2.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02y_bench_c5_1gpu_swap2.json"))
print("c5 swap=2 step %.2f ms kernel %.2f e2e %.2f warm %.2f frac %.3f parity %s launches %s" % (d["ms_per_step"], d["main_kernel_ms"], d["e2e"]["ms_per_step"], d["e2e_resident"]["ms_per_step"], d["roofline"]["frac"], d["parity"], d["gpu_launches"]))
PY

#!/bin/bash
# round 2, call W (1 GPU): swapped SEED launch (candidates as rows, reads as lanes) -- parity, phase times, c5 bench line,
# compute-sanitizer over the two-level 2-set path
TAG=r02w
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "two_level or similarity_order" > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log
ISOCON_NN_DEBUG=2 timeout 300 python tools/phase_times.py c5 1.0 > gpurun_out/${TAG}_phase_times_c5.txt 2>&1; grep -E "^rep|first cap|swapped" gpurun_out/${TAG}_phase_times_c5.txt | tail -5
timeout 600 python bench.py --workload c5 --steps 5 --warmup 3 --cpu-queries 64 2> gpurun_out/${TAG}_b5.err | grep '^{' > gpurun_out/${TAG}_bench_c5_1gpu.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02w_bench_c5_1gpu.json"))
print("c5 step %.2f ms kernel %.2f e2e %.2f warm %.2f frac %.3f parity %s launches %s" % (d["ms_per_step"], d["main_kernel_ms"], d["e2e"]["ms_per_step"], d["e2e_resident"]["ms_per_step"], d["roofline"]["frac"], d["parity"], d["gpu_launches"]))
PY
for tool in memcheck racecheck; do
  timeout 420 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize.py --c5 > gpurun_out/${TAG}_sanitizer_${tool}_c5.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize workload ok|Error|hazard" gpurun_out/${TAG}_sanitizer_${tool}_c5.log | head -6
done

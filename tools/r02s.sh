#!/bin/bash
# round 2, call S (1 GPU): first cap of the hinted SEED rows picked by a sample -- parity, phase times, c5 bench line, e2e profile
TAG=r02s
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "two_level or similarity_order" > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log
ISOCON_NN_DEBUG=2 timeout 300 python tools/phase_times.py c5 1.0 > gpurun_out/${TAG}_phase_times_c5.txt 2>&1; grep -E "^rep|first cap" gpurun_out/${TAG}_phase_times_c5.txt | tail -4
ISOCON_NN_SEED_SAMPLE=0 timeout 300 python tools/phase_times.py c5 1.0 2>&1 | grep -E "^rep" | tail -2
timeout 600 python bench.py --workload c5 --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/${TAG}_b5.err | grep '^{' > gpurun_out/${TAG}_bench_c5_1gpu.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02s_bench_c5_1gpu.json"))
print("c5 step %.2f ms kernel %.2f e2e %.2f warm %.2f parity %s launches %s" % (d["ms_per_step"], d["main_kernel_ms"], d["e2e"]["ms_per_step"], d["e2e_resident"]["ms_per_step"], d["parity"], d["gpu_launches"]))
PY
timeout 300 python tools/e2e_profile.py c5 1.0 > gpurun_out/${TAG}_e2e_profile_c5.txt 2>&1; grep -E "^==" gpurun_out/${TAG}_e2e_profile_c5.txt

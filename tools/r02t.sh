#!/bin/bash
# round 2, call T (1 GPU): level 1 by alignment on the later ladder passes, helper thread gated on the library call,
# dict addresses from prepare to fill -- parity, c5 and c2 bench lines, e2e profile
TAG=r02t
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "two_level or similarity_order or install or foreign or residency or upload or edge_buffer" > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log
ISOCON_NN_DEBUG=2 timeout 300 python tools/phase_times.py c5 1.0 > gpurun_out/${TAG}_phase_times_c5.txt 2>&1; grep -E "^rep|two-level pass" gpurun_out/${TAG}_phase_times_c5.txt | tail -4
timeout 600 python bench.py --workload c5 --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/${TAG}_b5.err | grep '^{' > gpurun_out/${TAG}_bench_c5_1gpu.json
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/${TAG}_b2.err | grep '^{' > gpurun_out/${TAG}_bench_c2_1gpu.json
python - <<'PY'
import json
for w in ("c5", "c2"):
    d = json.load(open("gpurun_out/r02t_bench_%s_1gpu.json" % w))
    print("%s step %.2f ms kernel %.2f e2e %.2f warm %.2f parity %s launches %s" % (w, d["ms_per_step"], d["main_kernel_ms"], d["e2e"]["ms_per_step"], d["e2e_resident"]["ms_per_step"], d["parity"], d["gpu_launches"]))
PY
timeout 300 python tools/e2e_profile.py c5 1.0 > gpurun_out/${TAG}_e2e_profile_c5.txt 2>&1; grep -E "^==" gpurun_out/${TAG}_e2e_profile_c5.txt

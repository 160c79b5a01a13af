#!/bin/bash
# round 2, call K (1 GPU): primer launch A/B, parity suite, host-side profiles (c2 e2e cold / warm, c5 phases)
TAG=r02k
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log
for pr in 0 1; do
  ISOCON_NN_PRIMER=$pr timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | grep '^{' > gpurun_out/${TAG}_c2_primer$pr.json
  ISOCON_NN_PRIMER=$pr timeout 600 python bench.py --workload c3 --scale 0.4 --steps 3 --warmup 1 --no-cpu-baseline 2>/dev/null | grep '^{' > gpurun_out/${TAG}_c3_primer$pr.json
done
python - <<'PY'
import json
for f in ("c2_primer0", "c2_primer1", "c3_primer0", "c3_primer1"):
    try:
        d = json.load(open("gpurun_out/r02k_%s.json" % f))
        st = d["device_stats"]
        print(f, "step %.2f ms kernel %.2f e2e %.2f warm %.2f | frac %.3f exec %.3f | wc %.4e parity %s" % (
            d["ms_per_step"], d["main_kernel_ms"], d["e2e"]["ms_per_step"], d["e2e_resident"]["ms_per_step"], d["roofline"]["frac"],
            d["roofline"]["executed_alu_ops_frac_of_peak"], st["word_columns"], d["parity"]))
    except Exception as e:
        print(f, "unreadable:", e)
PY
timeout 300 python tools/e2e_profile.py c2 1.0 > gpurun_out/${TAG}_e2e_profile_c2.txt 2>&1; grep -E "^==|tottime|^ +[0-9]" gpurun_out/${TAG}_e2e_profile_c2.txt | head -44
ISOCON_NN_DEBUG=2 timeout 300 python tools/phase_times.py c5 1.0 > gpurun_out/${TAG}_phase_times_c5.txt 2>&1; tail -22 gpurun_out/${TAG}_phase_times_c5.txt

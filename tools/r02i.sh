#!/bin/bash
# round 2, call I (1 GPU): parity suite, bench lines with full-size parity (c2, c3@0.4, c5), pilot split A/B
TAG=r02i
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; tail -4 gpurun_out/${TAG}_pytest.log
for br in 0 20 40; do
  ISOCON_NN_BRIDGE=$br timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | grep '^{' > gpurun_out/${TAG}_c2_bridge$br.json
done
timeout 600 python bench.py --steps 5 --warmup 3 2> gpurun_out/${TAG}_b2.err | grep '^{' > gpurun_out/${TAG}_bench_c2_1gpu.json
timeout 900 python bench.py --workload c3 --scale 0.4 --steps 3 --warmup 1 --cpu-queries 64 2> gpurun_out/${TAG}_b3.err | grep '^{' > gpurun_out/${TAG}_bench_c3_s0.4_1gpu.json
timeout 900 python bench.py --workload c5 --steps 3 --warmup 1 --cpu-queries 64 2> gpurun_out/${TAG}_b5.err | grep '^{' > gpurun_out/${TAG}_bench_c5_1gpu.json
python - <<'PY'
import json
for f in ("c2_bridge0", "c2_bridge20", "c2_bridge40", "bench_c2_1gpu", "bench_c3_s0.4_1gpu", "bench_c5_1gpu"):
    try:
        d = json.load(open("gpurun_out/r02i_%s.json" % f))
        st = d["device_stats"]
        print(f, "step %.2f ms kernel %.2f e2e %.2f warm %.2f | frac %.3f exec %.3f | pilot_rows %d clusters %d parity %s" % (
            d["ms_per_step"], d["main_kernel_ms"], d["e2e"]["ms_per_step"], d["e2e_resident"]["ms_per_step"], d["roofline"]["frac"],
            d["roofline"]["executed_alu_ops_frac_of_peak"], st["pilot_rows"], st["clusters"], d["parity"]))
    except Exception as e:
        print(f, "unreadable:", e)
PY

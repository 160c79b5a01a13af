#!/bin/bash
# round 2, call O (1 GPU): min-hash similarity order of the targets of one-sided passes: parity + c5 A/B
TAG=r02o
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log
for cl in 1; do
  ISOCON_NN_CLUSTER=$cl timeout 900 python bench.py --workload c5 --steps 3 --warmup 1 --no-cpu-baseline 2>gpurun_out/${TAG}_c5_$cl.err | grep '^{' > gpurun_out/${TAG}_c5_cluster$cl.json
done
ISOCON_NN_QGRAM=0 timeout 900 python bench.py --workload c5 --steps 3 --warmup 1 --no-cpu-baseline 2>/dev/null | grep '^{' > gpurun_out/${TAG}_c5_hints_only.json
ISOCON_NN_DEBUG=2 timeout 300 python tools/phase_times.py c5 1.0 > gpurun_out/${TAG}_phase_times_c5.txt 2>&1; tail -16 gpurun_out/${TAG}_phase_times_c5.txt
python - <<'PY'
import json
for f in ("c5_cluster1", "c5_hints_only"):
    try:
        d = json.load(open("gpurun_out/r02o_%s.json" % f))
        st = d["device_stats"]
        print(f, "step %.2f ms kernel %.2f e2e %.2f warm %.2f | frac %.3f exec %.3f | wc %.4e cols %.4e clusters %d parity %s" % (
            d["ms_per_step"], d["main_kernel_ms"], d["e2e"]["ms_per_step"], d["e2e_resident"]["ms_per_step"], d["roofline"]["frac"],
            d["roofline"]["executed_alu_ops_frac_of_peak"], st["word_columns"], st["columns"], st["clusters"], d["parity"]))
    except Exception as e:
        print(f, "unreadable:", e)
PY
tail -3 gpurun_out/${TAG}_c5_1.err

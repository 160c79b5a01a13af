#!/bin/bash
# round 2, call D: similarity order of the MAIN targets -- parity + A/B on c2 and c3
TAG=r02d
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; tail -4 gpurun_out/${TAG}_pytest.log
for cl in 0 1; do
  ISOCON_NN_CLUSTER=$cl timeout 600 python bench.py --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_c2_cluster$cl.json 2> gpurun_out/${TAG}_c2_cluster$cl.err
  ISOCON_NN_CLUSTER=$cl timeout 900 python bench.py --workload c3 --scale 0.4 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_c3_cluster$cl.json 2> gpurun_out/${TAG}_c3_cluster$cl.err
done
python - <<'PY'
import json
for f in ("c2_cluster0", "c2_cluster1", "c3_cluster0", "c3_cluster1"):
    try:
        d = json.load(open("gpurun_out/r02d_%s.json" % f))
        st = d["device_stats"]
        print(f, "step %.2f ms kernel %.2f e2e %.2f | frac %.3f exec %.3f | wc %.3e cols %.3e clusters %d bins %d parity %s" % (
            d["ms_per_step"], d["main_kernel_ms"], d["e2e"]["ms_per_step"], d["roofline"]["frac"],
            d["roofline"]["executed_alu_ops_frac_of_peak"], st["word_columns"], st["columns"], st["clusters"], st["bins"], d["parity"]))
    except Exception as e:
        print(f, "unreadable:", e)
PY

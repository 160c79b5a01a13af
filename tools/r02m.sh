#!/bin/bash
# round 2, call M (1 GPU): A/B of layout knobs (minor sort key inside a cluster, pilot fraction, class width) on c2 and c3@0.4
TAG=r02m
mkdir -p gpurun_out
run() {  # name env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 3 --warmup 1 --no-cpu-baseline 2>/dev/null | grep '^{' > gpurun_out/${TAG}_c2_$name.json
  env "$@" timeout 600 python bench.py --workload c3 --scale 0.4 --steps 2 --warmup 1 --no-cpu-baseline 2>/dev/null | grep '^{' > gpurun_out/${TAG}_c3_$name.json
}
run base ISOCON_NN_ORDER_BEST=0
run best1 ISOCON_NN_ORDER_BEST=1
run best2 ISOCON_NN_ORDER_BEST=2
run div8 ISOCON_NN_PILOT_DIV=8
run div14 ISOCON_NN_PILOT_DIV=14
run div20 ISOCON_NN_PILOT_DIV=20
run gran16 ISOCON_NN_CLASS_GRAN=16
python - <<'PY'
import json
for name in ("base", "best1", "best2", "div8", "div14", "div20", "gran16"):
    row = [name]
    for w in ("c2", "c3"):
        try:
            d = json.load(open("gpurun_out/r02m_%s_%s.json" % (w, name)))
            row.append("%s step %.2f kernel %.2f frac %.3f exec %.3f wc %.4e parity %s" % (
                w, d["ms_per_step"], d["main_kernel_ms"], d["roofline"]["frac"], d["roofline"]["executed_alu_ops_frac_of_peak"],
                d["device_stats"]["word_columns"], d["parity"]["full_size_graph_digest_equals_oracle"]))
        except Exception as e:
            row.append("%s unreadable: %s" % (w, e))
    print(" | ".join(row))
PY

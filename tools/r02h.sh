#!/bin/bash
# round 2, call H (1 GPU): ncu captures of the shipped binary (c2: launch list + full capture; c4 slice: full capture),
# A/B of coarser window-width sets on c3 / c4 / c2
TAG=r02h
mkdir -p gpurun_out
for v in base coarse8 coarse6; do
  LIB=""; [ $v != base ] && LIB="$PWD/build/variants/libisocon_nn_$v.so"
  for w in "c4 0.25" "c3 0.4" "c2 1.0"; do
    set -- $w
    ISOCON_NN_LIB=$LIB timeout 600 python bench.py --workload $1 --scale $2 --steps 2 --warmup 1 --no-cpu-baseline --no-parity 2>/dev/null | grep '^{' > gpurun_out/${TAG}_${v}_$1.json
  done
done
python - <<'PY'
import json
for v in ("base", "coarse8", "coarse6"):
    for w in ("c4", "c3", "c2"):
        try:
            d = json.load(open("gpurun_out/r02h_%s_%s.json" % (v, w)))
            print(v, w, "step %.2f ms kernel %.2f | frac %.3f exec %.3f" % (d["ms_per_step"], d["main_kernel_ms"], d["roofline"]["frac"], d["roofline"]["executed_alu_ops_frac_of_peak"]))
        except Exception as e:
            print(v, w, "unreadable:", e)
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_bench_c2.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:nn_row_kernel -s 2 -c 2 -f -o gpurun_out/prof_row_${TAG}_c2 \
    python tools/phase_times.py c2 1.0 > gpurun_out/${TAG}_ncu_full_c2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:nn_row_kernel -s 2 -c 2 -f -o gpurun_out/prof_row_${TAG}_c4 \
    python tools/phase_times.py c4 0.1 > gpurun_out/${TAG}_ncu_full_c4.log 2>&1
ls -la gpurun_out | grep ${TAG} | tail -20

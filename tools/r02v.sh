#!/bin/bash
# round 2, call V (1 GPU): the final set -- parity suite, bench lines of every BASELINE config with its full-size digest,
# reference arm, ncu launch list and full captures (row kernel on c2, tile kernel on c5) of the final binary
TAG=r02v
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py 2> gpurun_out/${TAG}_b2.err | grep '^{' > gpurun_out/${TAG}_bench_c2_1gpu.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | grep '^{' > gpurun_out/${TAG}_bench_c2_reference_arm.json
timeout 900 python bench.py --workload c4 --scale 0.25 --steps 2 --warmup 1 --no-cpu-baseline 2> gpurun_out/${TAG}_b4.err | grep '^{' > gpurun_out/${TAG}_bench_c4_s0.25_1gpu.json
timeout 900 python bench.py --workload c3 --scale 0.4 --steps 3 --warmup 1 --no-cpu-baseline 2> gpurun_out/${TAG}_b3.err | grep '^{' > gpurun_out/${TAG}_bench_c3_s0.4_1gpu.json
timeout 900 python bench.py --workload c5 --steps 5 --warmup 3 --cpu-queries 64 2> gpurun_out/${TAG}_b5.err | grep '^{' > gpurun_out/${TAG}_bench_c5_1gpu.json
python - <<'PY'
import json
for f in ("bench_c2_1gpu", "bench_c4_s0.25_1gpu", "bench_c3_s0.4_1gpu", "bench_c5_1gpu"):
    try:
        d = json.load(open("gpurun_out/r02v_%s.json" % f))
        st = d["device_stats"]
        print(f, "step %.2f ms kernel %.2f e2e %.2f warm %.2f | frac %.3f exec %.3f | clusters %d parity %s cpu %s" % (
            d["ms_per_step"], d["main_kernel_ms"], d["e2e"]["ms_per_step"], d["e2e_resident"]["ms_per_step"], d["roofline"]["frac"],
            d["roofline"]["executed_alu_ops_frac_of_peak"], st["clusters"], d["parity"], (d.get("cpu_baseline") or {}).get("value")))
    except Exception as e:
        print(f, "unreadable:", e)
try:
    d = json.load(open("gpurun_out/r02v_bench_c2_reference_arm.json"))
    print("reference arm:", d["value"], d["ms_per_step"], d["cpu_baseline"]["sample"], d["cpu_baseline"].get("sample_fraction"))
except Exception as e:
    print("reference arm unreadable:", e)
PY
timeout 300 python tools/e2e_profile.py c2 1.0 > gpurun_out/${TAG}_e2e_profile_c2.txt 2>&1; grep -E "^==" gpurun_out/${TAG}_e2e_profile_c2.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_bench_c2.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:nn_row_kernel -s 2 -c 2 -f -o gpurun_out/prof_row_${TAG}_c2 \
    python tools/phase_times.py c2 1.0 > gpurun_out/${TAG}_ncu_full_c2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:nn_tile_kernel -s 1 -c 1 -f -o gpurun_out/prof_tile_${TAG}_c5 \
    python tools/phase_times.py c5 1.0 > gpurun_out/${TAG}_ncu_full_c5.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_graph_c5.csv \
    python tools/phase_times.py c5 1.0 > gpurun_out/${TAG}_c5_under_ncu.log 2>&1
ls -la gpurun_out | grep ${TAG} | tail -18

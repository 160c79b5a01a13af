#!/bin/bash
# round 2, call E (2 GPUs): sharded-vs-alone parity incl. similarity order / foreign reads / regrowth / finite depth,
# 2-GPU bench lines of c2, c3 (0.4), c5, and the 1-GPU c2 line of the same binary
TAG=r02e
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29511 tools/check_multi_gpu.py > gpurun_out/${TAG}_check_2gpu.log 2>&1; echo "check rc=$?"; grep "^rank" gpurun_out/${TAG}_check_2gpu.log | sort | tail -30
timeout 600 python -m pytest tests -m gpu -x -q -k "two_gpus or similarity" > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_c2_1gpu.json 2> gpurun_out/${TAG}_b1.err; tail -c 300 gpurun_out/${TAG}_b1.err
timeout 600 $TR --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_c2_2gpu.json 2> gpurun_out/${TAG}_b2.err; tail -c 300 gpurun_out/${TAG}_b2.err
timeout 600 $TR --master-port 29513 bench.py --gpus 2 --workload c5 --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_c5_2gpu.json 2> gpurun_out/${TAG}_b5.err; tail -c 300 gpurun_out/${TAG}_b5.err
timeout 600 $TR --master-port 29514 bench.py --gpus 2 --workload c3 --scale 0.4 --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_c3_2gpu.json 2> gpurun_out/${TAG}_b3.err; tail -c 300 gpurun_out/${TAG}_b3.err
python - <<'PY'
import json
for f in ("bench_c2_1gpu", "bench_c2_2gpu", "bench_c5_2gpu", "bench_c3_2gpu"):
    try:
        d = json.load(open("gpurun_out/r02e_%s.json" % f))
        st = d["device_stats"]
        print(f, "step %.2f ms kernel %.2f e2e %.2f warm %.2f | frac %.3f exec %.3f | clusters %d parity %s" % (
            d["ms_per_step"], d["main_kernel_ms"], d["e2e"]["ms_per_step"], d["e2e_resident"]["ms_per_step"], d["roofline"]["frac"],
            d["roofline"]["executed_alu_ops_frac_of_peak"], st["clusters"], d["parity"]))
        if "sharding_rank0_last_step" in d:
            print("   ", d["sharding_rank0_last_step"])
    except Exception as e:
        print(f, "unreadable:", e)
PY

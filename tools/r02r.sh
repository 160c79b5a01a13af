#!/bin/bash
# round 2, call R (8 GPUs): final binary -- sharded-vs-alone at 8 ranks, bench lines of c2 / c5 / c3@0.4 / c4@0.25
TAG=r02r
N=8
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29521 tools/check_multi_gpu.py > gpurun_out/${TAG}_check_${N}gpu.log 2>&1
echo "check rc=$? same=$(grep -c -- '-> same' gpurun_out/${TAG}_check_${N}gpu.log) different=$(grep -c DIFFERENT gpurun_out/${TAG}_check_${N}gpu.log)"
timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/${TAG}_b2.err | grep '^{' > gpurun_out/${TAG}_bench_c2_${N}gpu.json
timeout 600 $TR --master-port 29513 bench.py --gpus $N --workload c5 --steps 3 --warmup 1 --no-cpu-baseline 2> gpurun_out/${TAG}_b5.err | grep '^{' > gpurun_out/${TAG}_bench_c5_${N}gpu.json
timeout 600 $TR --master-port 29514 bench.py --gpus $N --workload c3 --scale 0.4 --steps 3 --warmup 1 --no-cpu-baseline 2> gpurun_out/${TAG}_b3.err | grep '^{' > gpurun_out/${TAG}_bench_c3_s0.4_${N}gpu.json
timeout 600 $TR --master-port 29515 bench.py --gpus $N --workload c4 --scale 0.25 --steps 2 --warmup 1 --no-cpu-baseline 2> gpurun_out/${TAG}_b4.err | grep '^{' > gpurun_out/${TAG}_bench_c4_s0.25_${N}gpu.json
python - <<PY
import json
for f in ("bench_c2_${N}gpu", "bench_c5_${N}gpu", "bench_c3_s0.4_${N}gpu", "bench_c4_s0.25_${N}gpu"):
    try:
        d = json.load(open("gpurun_out/${TAG}_%s.json" % f))
        print(f, "step %.2f ms kernel %.2f e2e %.2f warm %.2f | frac %.3f | parity %s" % (
            d["ms_per_step"], d["main_kernel_ms"], d["e2e"]["ms_per_step"], d["e2e_resident"]["ms_per_step"], d["roofline"]["frac"], d["parity"]))
    except Exception as e:
        print(f, "unreadable:", e)
PY

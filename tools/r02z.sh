#!/bin/bash
# round 2, call Z (8 GPUs): final binary -- bench lines c5 and c2
TAG=r02z
N=8
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 60 $TR --master-port 29571 bench.py --gpus $N --workload c5 --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/${TAG}_b5.err | grep '^{' > gpurun_out/${TAG}_bench_c5_${N}gpu.json
timeout 40 $TR --master-port 29572 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/${TAG}_b2.err | grep '^{' > gpurun_out/${TAG}_bench_c2_${N}gpu.json
python - <<PY
import json
for f in ("bench_c5_${N}gpu", "bench_c2_${N}gpu"):
    try:
        d = json.load(open("gpurun_out/${TAG}_%s.json" % f))
        print(f, "step %.2f ms kernel %.2f e2e %.2f warm %.2f | parity %s" % (
            d["ms_per_step"], d["main_kernel_ms"], d["e2e"]["ms_per_step"], d["e2e_resident"]["ms_per_step"], d["parity"]))
    except Exception as e:
        print(f, "unreadable:", e)
PY

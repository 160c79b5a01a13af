#!/bin/bash
# round 2, call Q (2 GPUs): the two-level one-sided pass across ranks: sharded-vs-alone (fused and collective), c5 / c2 lines
TAG=r02q
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for i in 1 2 3; do
  FU=1; [ $i -gt 2 ] && FU=0
  ISOCON_NN_FUSE=$FU timeout 600 $TR --master-port $((29520 + i)) tools/check_multi_gpu.py > gpurun_out/${TAG}_check_$i.log 2>&1
  echo "run $i fuse=$FU rc=$? same=$(grep -c -- '-> same' gpurun_out/${TAG}_check_$i.log) different=$(grep -c DIFFERENT gpurun_out/${TAG}_check_$i.log)"
  grep -E "DIFFERENT|Error|error" gpurun_out/${TAG}_check_$i.log | head -5
done
timeout 600 $TR --master-port 29513 bench.py --gpus 2 --workload c5 --steps 3 --warmup 1 --no-cpu-baseline 2> gpurun_out/${TAG}_b5.err | grep '^{' > gpurun_out/${TAG}_bench_c5_2gpu.json; tail -c 300 gpurun_out/${TAG}_b5.err
ISOCON_NN_FUSE=0 timeout 600 $TR --master-port 29516 bench.py --gpus 2 --workload c5 --steps 3 --warmup 1 --no-cpu-baseline 2>/dev/null | grep '^{' > gpurun_out/${TAG}_bench_c5_2gpu_nofuse.json
timeout 600 $TR --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | grep '^{' > gpurun_out/${TAG}_bench_c2_2gpu.json
python - <<'PY'
import json
for f in ("bench_c5_2gpu", "bench_c5_2gpu_nofuse", "bench_c2_2gpu"):
    try:
        d = json.load(open("gpurun_out/r02q_%s.json" % f))
        print(f, "step %.2f ms kernel %.2f e2e %.2f warm %.2f | parity %s" % (
            d["ms_per_step"], d["main_kernel_ms"], d["e2e"]["ms_per_step"], d["e2e_resident"]["ms_per_step"], d["parity"]))
    except Exception as e:
        print(f, "unreadable:", e)
PY

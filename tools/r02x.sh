#!/bin/bash
# round 2, call X (2 GPUs): swapped SEED launch under two ranks -- sharded vs alone (fused, fused with level 2 swapped,
# through the collectives), bench lines c5 / c2
TAG=r02x
N=2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for i in 1 2 3; do
  FUSE=1; SWAP=1; [ $i = 2 ] && SWAP=2; [ $i = 3 ] && FUSE=0
  ISOCON_NN_SWAP=$SWAP ISOCON_NN_FUSE=$FUSE timeout 600 $TR --master-port $((29550 + i)) tools/check_multi_gpu.py > gpurun_out/${TAG}_check_$i.log 2>&1
  echo "run $i fuse=$FUSE swap=$SWAP rc=$? same=$(grep -c -- '-> same' gpurun_out/${TAG}_check_$i.log) different=$(grep -c DIFFERENT gpurun_out/${TAG}_check_$i.log)"
done
timeout 600 $TR --master-port 29561 bench.py --gpus $N --workload c5 --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/${TAG}_b5.err | grep '^{' > gpurun_out/${TAG}_bench_c5_${N}gpu.json
timeout 600 $TR --master-port 29562 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/${TAG}_b2.err | grep '^{' > gpurun_out/${TAG}_bench_c2_${N}gpu.json
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

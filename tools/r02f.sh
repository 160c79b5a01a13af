#!/bin/bash
# round 2, call F (2 GPUs): repeat the sharded-vs-alone check to catch a rare difference
TAG=r02f
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for i in 1 2 3 4 5 6; do
  CL=1; [ $i -gt 4 ] && CL=0
  ISOCON_NN_CLUSTER=$CL timeout 600 $TR --master-port $((29520 + i)) tools/check_multi_gpu.py > gpurun_out/${TAG}_check_$i.log 2>&1
  echo "run $i cluster=$CL rc=$? same=$(grep -c -- '-> same' gpurun_out/${TAG}_check_$i.log) different=$(grep -c DIFFERENT gpurun_out/${TAG}_check_$i.log)"
  grep DIFFERENT gpurun_out/${TAG}_check_$i.log
done

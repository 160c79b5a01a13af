#!/bin/bash
# 2 GPUs of one box: sharded-vs-alone parity (and optionally the c2 and c5 bench lines: run_2gpu.sh bench).
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$TR --master-port 29511 tools/check_multi_gpu.py > gpurun_out/check_2gpu.log 2>&1; echo "check rc=$?"; grep "^rank" gpurun_out/check_2gpu.log | sort | tail -12
if [ "$1" = "bench" ]; then
$TR --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_c2_2gpu.json 2> gpurun_out/b2.err; tail -c 300 gpurun_out/b2.err
$TR --master-port 29513 bench.py --gpus 2 --workload c5 --steps 2 --warmup 1 > gpurun_out/bench_c5_2gpu.json 2> gpurun_out/b5.err; tail -c 300 gpurun_out/b5.err
fi

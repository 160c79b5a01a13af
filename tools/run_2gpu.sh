#!/bin/bash
# 2 GPUs of one box: sharded-vs-alone parity, then the c2 and c5 bench lines.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$TR --master-port 29511 tools/check_multi_gpu.py > gpurun_out/r01e_check_2gpu.log 2>&1; echo "check rc=$?"; grep "rank" gpurun_out/r01e_check_2gpu.log | tail -12
$TR --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r01e_bench_c2_2gpu.json 2> gpurun_out/b2.err; tail -c 300 gpurun_out/b2.err
$TR --master-port 29513 bench.py --gpus 2 --workload c5 --steps 2 --warmup 1 > gpurun_out/r01e_bench_c5_2gpu.json 2> gpurun_out/b5.err; tail -c 300 gpurun_out/b5.err

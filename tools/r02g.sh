#!/bin/bash
# round 2, call G (2 GPUs): fused multi-rank flow (device barriers, no collectives) vs collective flow: repeated
# sharded-vs-alone checks, then bench lines
TAG=r02g
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for i in 1 2 3 4 5 6; do
  FU=1; [ $i -gt 4 ] && FU=0
  ISOCON_NN_FUSE=$FU timeout 600 $TR --master-port $((29520 + i)) tools/check_multi_gpu.py > gpurun_out/${TAG}_check_$i.log 2>&1
  echo "run $i fuse=$FU rc=$? same=$(grep -c -- '-> same' gpurun_out/${TAG}_check_$i.log) different=$(grep -c DIFFERENT gpurun_out/${TAG}_check_$i.log)"
  grep -E "DIFFERENT|Error|error" gpurun_out/${TAG}_check_$i.log | head -5
done
timeout 600 $TR --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 2> gpurun_out/${TAG}_b2.err | grep '^{' > gpurun_out/${TAG}_bench_c2_2gpu.json; tail -c 300 gpurun_out/${TAG}_b2.err
ISOCON_NN_FUSE=0 timeout 600 $TR --master-port 29515 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/${TAG}_b2n.err | grep '^{' > gpurun_out/${TAG}_bench_c2_2gpu_nofuse.json
timeout 600 $TR --master-port 29513 bench.py --gpus 2 --workload c5 --steps 2 --warmup 1 2> gpurun_out/${TAG}_b5.err | grep '^{' > gpurun_out/${TAG}_bench_c5_2gpu.json; tail -c 300 gpurun_out/${TAG}_b5.err
timeout 600 $TR --master-port 29514 bench.py --gpus 2 --workload c3 --scale 0.4 --steps 2 --warmup 1 2> gpurun_out/${TAG}_b3.err | grep '^{' > gpurun_out/${TAG}_bench_c3_2gpu.json; tail -c 300 gpurun_out/${TAG}_b3.err
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_c2_1gpu.json 2> gpurun_out/${TAG}_b1.err
python - <<'PY'
import json
for f in ("bench_c2_1gpu", "bench_c2_2gpu", "bench_c2_2gpu_nofuse", "bench_c5_2gpu", "bench_c3_2gpu"):
    try:
        d = json.load(open("gpurun_out/r02g_%s.json" % f))
        st = d["device_stats"]
        print(f, "step %.2f ms kernel %.2f e2e %.2f warm %.2f | frac %.3f exec %.3f | clusters %d parity %s" % (
            d["ms_per_step"], d["main_kernel_ms"], d["e2e"]["ms_per_step"], d["e2e_resident"]["ms_per_step"], d["roofline"]["frac"],
            d["roofline"]["executed_alu_ops_frac_of_peak"], st["clusters"], d["parity"]))
        if "sharding_rank0_last_step" in d:
            print("   ", d["sharding_rank0_last_step"])
    except Exception as e:
        print(f, "unreadable:", e)
PY

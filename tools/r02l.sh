#!/bin/bash
# round 2, call L (8 GPUs): primer launch A/B at 8 ranks (c2, c3@0.4), 4-rank c2 line
TAG=r02l
mkdir -p gpurun_out
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
for pr in 0 4; do
  ISOCON_NN_PRIMER=$pr timeout 300 $TR8 --master-port $((29530 + pr)) bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | grep '^{' > gpurun_out/${TAG}_c2_8gpu_primer$pr.json
  ISOCON_NN_PRIMER=$pr timeout 300 $TR8 --master-port $((29540 + pr)) bench.py --gpus 8 --workload c3 --scale 0.4 --steps 3 --warmup 1 --no-cpu-baseline 2>/dev/null | grep '^{' > gpurun_out/${TAG}_c3_8gpu_primer$pr.json
done
timeout 300 $TR4 --master-port 29550 bench.py --gpus 4 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | grep '^{' > gpurun_out/${TAG}_bench_c2_4gpu.json
timeout 300 $TR4 --master-port 29551 bench.py --gpus 4 --workload c5 --steps 3 --warmup 1 --no-cpu-baseline 2>/dev/null | grep '^{' > gpurun_out/${TAG}_bench_c5_4gpu.json
python - <<'PY'
import json
for f in ("c2_8gpu_primer0", "c2_8gpu_primer4", "c3_8gpu_primer0", "c3_8gpu_primer4", "bench_c2_4gpu", "bench_c5_4gpu"):
    try:
        d = json.load(open("gpurun_out/r02l_%s.json" % f))
        st = d["device_stats"]
        print(f, "step %.2f ms kernel %.2f e2e %.2f warm %.2f | frac %.3f exec %.3f | wc %.4e parity %s" % (
            d["ms_per_step"], d["main_kernel_ms"], d["e2e"]["ms_per_step"], d["e2e_resident"]["ms_per_step"], d["roofline"]["frac"],
            d["roofline"]["executed_alu_ops_frac_of_peak"], st["word_columns"], d["parity"]))
    except Exception as e:
        print(f, "unreadable:", e)
PY

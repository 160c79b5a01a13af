#!/bin/bash
# One B200 call: GPU parity suite, bench line, reference arm, ncu launch list of the bench, ncu full capture of the
# row kernel (second graph build of tools/phase_times.py: launch 0 = PILOT, launch 1 = MAIN).   usage: profile_round.sh r01e
TAG=${1:-r01x}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log
python bench.py > gpurun_out/${TAG}_bench_c2_1gpu.json 2> gpurun_out/${TAG}_bench_c2_1gpu.err; tail -c 300 gpurun_out/${TAG}_bench_c2_1gpu.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_c2_reference_arm.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_bench_c2.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:nn_row_kernel -s 2 -c 2 -f -o gpurun_out/prof_row_${TAG} \
    python tools/phase_times.py c2 1.0 > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out | tail -12

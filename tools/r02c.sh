#!/bin/bash
# round 2, call C: compute-sanitizer over the kernels (memcheck, racecheck, synccheck, initcheck) + the parity suite
TAG=r02c
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log
for tool in memcheck racecheck synccheck initcheck; do
  extra="--quick"; [ $tool = memcheck ] && extra=""
  timeout 420 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize.py $extra > gpurun_out/${TAG}_sanitizer_${tool}.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize workload ok|Error|hazard" gpurun_out/${TAG}_sanitizer_${tool}.log | head -8
done

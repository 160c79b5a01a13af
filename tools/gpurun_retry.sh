#!/bin/bash
# usage: tools/gpurun_retry.sh <gpus> <timeout> <command...>   -- retries while the pod answers "busy" (nothing charged)
G=$1; T=$2; shift 2
for attempt in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient"; then sleep 90; continue; fi
  echo "$out"; exit $rc
done
echo "gave up after 40 busy answers"; exit 3

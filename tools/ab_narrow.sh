#!/bin/bash
# A/B of the shrinking window on a B200: ISOCON_NN_NARROW = check interval in 32-column chunks (0 = off).
mkdir -p gpurun_out
for N in 0 1 2 3 4 6 8; do
  ISOCON_NN_NARROW=$N python tools/phase_times.py c2 1.0 > gpurun_out/ab_narrow_c2_$N.log 2>&1
done
for N in 0 2 4; do
  ISOCON_NN_NARROW=$N python tools/phase_times.py c3 0.2 > gpurun_out/ab_narrow_c3_$N.log 2>&1
  ISOCON_NN_NARROW=$N python tools/phase_times.py c4 0.1 > gpurun_out/ab_narrow_c4_$N.log 2>&1
done
tail -n 3 gpurun_out/ab_narrow_*.log

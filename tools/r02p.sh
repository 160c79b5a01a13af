#!/bin/bash
# round 2, call P (1 GPU): where the fixed cost of the two-level c5 graph goes
TAG=r02p
mkdir -p gpurun_out
ISOCON_NN_DEBUG=2 timeout 300 python tools/phase_times.py c5 1.0 > gpurun_out/${TAG}_phase_times_c5.txt 2>&1; tail -42 gpurun_out/${TAG}_phase_times_c5.txt

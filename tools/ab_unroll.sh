#!/bin/bash
# A/B of the unroll policy of the diagonal-band chunk (build/libisocon_nn_u*.so; u0 = by width) x check interval.
mkdir -p gpurun_out
run() { ISOCON_NN_LIB=$PWD/build/libisocon_nn_u$1.so ISOCON_NN_NARROW=$2 python tools/phase_times.py $3 $4 2>&1 | tail -2 | head -1 | sed "s/^/$3 U=$1 N=$2 /"; }
for U in 0 2 1; do run $U 4 c2 1.0; run $U 4 c4 0.1; run $U 4 c3 0.2; done
for N in 2 3 6 8; do run 0 $N c2 1.0; run 0 $N c4 0.1; done
run 0 0 c3 0.2

#!/bin/bash
# A/B on one B200: unroll policies (build/libisocon_nn_p{1,2}.so vs the shipped one).
run() { env $1 python tools/phase_times.py $2 $3 2>&1 | tail -2 | head -1 | sed "s|^|$2 $1 |"; }
for W in "c2 1.0" "c4 0.1" "c3 0.2"; do
  run "X=0" $W
  run "ISOCON_NN_LIB=$PWD/build/libisocon_nn_p1.so" $W
  run "ISOCON_NN_LIB=$PWD/build/libisocon_nn_p2.so" $W
done

#!/bin/bash
# A/B two builds of librtb200.so on the same GPU in one gpurun call: tools/ab_libs.sh <what> (explore.py mode)
for rep in 1 2; do
  for lib in librtb200_prev.so librtb200.so; do
    echo "== $lib (rep $rep)"
    RTB200_LIB=$PWD/ray_tracing_series_rust_b200/$lib bash tools/q.sh ${1:-all} | grep -v COUNT | grep -v warm
  done
done

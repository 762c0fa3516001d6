#!/bin/bash
# A/B several builds of librtb200.so on the same GPU in one gpurun call:
#   tools/ab_libs.sh <explore.py mode> [lib suffixes...]     default: prev and the current build
what=${1:-all}; shift
libs=${@:-"_prev "}
for rep in 1 2; do
  for sfx in $libs ""; do
    lib=librtb200$sfx.so
    echo "== $lib (rep $rep)"
    RTB200_LIB=$PWD/ray_tracing_series_rust_b200/$lib bash tools/q.sh $what | grep -v COUNT | grep -v warm
  done
done

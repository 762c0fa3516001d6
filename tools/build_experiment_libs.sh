#!/bin/bash
# Builds the compile-time experiment variants of librtb200.so next to the product library (here, on the CPU box: nvcc
# cross-compiles; the .so files travel to the GPU box with the snapshot).  See DESIGN.md section 9.
set -e
cd "$(dirname "$0")/../ray_tracing_series_rust_b200/csrc"
make -j1 OUT=../librtb200_near.so B=build_near EXTRA=-DRT_WIDE_NEAREST_ONLY > /dev/null
make -j1 OUT=../librtb200_pf1.so B=build_pf1 EXTRA=-DRT_WIDE_PREFETCH=1 > /dev/null
make -j1 OUT=../librtb200_pf2.so B=build_pf2 EXTRA=-DRT_WIDE_PREFETCH=2 > /dev/null
make -j1 OUT=../librtb200_tile.so B=build_tile EXTRA=-DRT_TILE_ORDER > /dev/null
ls -la ../librtb200*.so

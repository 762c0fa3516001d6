#!/bin/bash
# One gpurun call that measures everything DESIGN.md section 9 lists as prepared-but-unmeasured (about 3 GPU-minutes).
#   tools/build_experiment_libs.sh                      (on the CPU box first)
#   gpurun --timeout 300 -- 'bash tools/round2_ab.sh > gpurun_out/round2_ab.log 2>&1'
set -u
echo "== fused: pairs vs 4-wide (identity + timing)";            timeout 60 python tools/ab_wide.py
echo "== wavefront kernels on the 4-wide collapse (forced)";     timeout 60 python tools/ab_wide.py wave
echo "== motion form of the wide nodes (book-1 as shipped)";     timeout 40 python tools/ab_wide.py moving
echo "== 6 CTAs/SM for the wide book-1 kernel"
for occ in 5 6 5 6; do RTB200_WIDE_OCC=$occ timeout 30 bash tools/q.sh fused | grep "book1 final" | sed "s/^/occ $occ: /"; done
echo "== compile-time variants (product library last)"
for rep in 1 2; do
  for sfx in _near _pf1 _pf2 _tile ""; do
    lib=$PWD/ray_tracing_series_rust_b200/librtb200$sfx.so
    [ -f "$lib" ] || continue
    RTB200_LIB=$lib RTB200_TILE_ORDER=1 timeout 40 bash tools/q.sh fused | grep -v warm | sed "s/^/lib$sfx (rep $rep): /"
  done
done

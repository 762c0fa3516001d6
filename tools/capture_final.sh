#!/bin/bash
# Final round-2 ncu evidence (kernels after the signed-row walk, multi-draw samplers, merged rejection loop, camera batches) for profiles/ (run under gpurun, one GPU).  Every ncu command follows a plain run of the same command
# that exited 0.  Outputs: gpurun_out/r2_*.csv (launch lists), gpurun_out/prof_r2_*.{raw,source}.csv
set -u
O=gpurun_out
run() { python tools/profile_target.py "$@"; }
B="python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline"
# 0. the launch list of the bench command itself (kernel share of a step)
$B > $O/plain_bench.log 2>&1 && timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_60_launches_bench_book1.csv $B > $O/ncu_bench.log 2>&1
# 1. book-1 final scene (RT_MODE_AUTO = fused, 4-wide walk): full set of k_mega
run book1 50 > $O/plain_b1.log 2>&1 && timeout 200 ncu --set full --clock-control none --import-source on -k regex:'^k_mega(_r)?$' -c 1 -o $O/prof_r2_mega_book1 python tools/profile_target.py book1 50 > $O/ncu_b1.log 2>&1
# 2. 871 200-triangle mesh room: k_mega_r
run mesh 2 > $O/plain_me.log 2>&1 && timeout 200 ncu --set full --clock-control none --import-source on -k regex:'^k_mega(_r)?$' -c 1 -o $O/prof_r2_mega_mesh python tools/profile_target.py mesh 2 > $O/ncu_me.log 2>&1
# 3. book-2 final scene, wavefront: launch list + full set of k_extend / k_shade_all
run book2 8 > $O/plain_b2.log 2>&1 && timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file $O/r2_60_launches_book2_8spp.csv python tools/profile_target.py book2 8 > $O/ncu_b2a.log 2>&1
run book2 8 > $O/plain_b2.log 2>&1 && timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_extend|k_shade_all' -s 20 -c 4 -o $O/prof_r2_wave_book2 python tools/profile_target.py book2 8 > $O/ncu_b2b.log 2>&1
# 4. Cornell smoke, wavefront
run smoke 20 > $O/plain_sm.log 2>&1 && timeout 200 ncu --set full --clock-control none --import-source on -k regex:'k_extend|k_shade_all' -s 20 -c 2 -o $O/prof_r2_wave_smoke python tools/profile_target.py smoke 20 > $O/ncu_sm.log 2>&1
tail -n 1 $O/plain_b1.log $O/plain_b2.log $O/plain_sm.log $O/plain_me.log
for r in prof_r2_mega_book1 prof_r2_mega_mesh prof_r2_wave_book2 prof_r2_wave_smoke; do
  [ -f $O/$r.ncu-rep ] || continue
  ncu -i $O/$r.ncu-rep --page raw --csv > $O/$r.raw.csv 2>/dev/null
  ncu -i $O/$r.ncu-rep --page source --csv > $O/$r.source.csv 2>/dev/null
  rm -f $O/$r.ncu-rep
done
ls -la $O | tail -30

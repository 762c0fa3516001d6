#!/bin/bash
# ncu evidence for profiles/ (run under gpurun, one GPU).  Every ncu command follows a plain run of the same command.
set -u
O=gpurun_out
run() { python tools/profile_target.py "$@"; }
# 1. book-1 final scene, fused mode (RT_MODE_AUTO): launch list + full set of k_mega
run book1 50 > $O/plain_b1.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 50 --csv --log-file $O/launches_book1.csv python tools/profile_target.py book1 50 > $O/ncu_b1a.log 2>&1
run book1 50 > $O/plain_b1.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'^k_mega(_r)?$' -c 1 -o $O/prof_mega_book1 python tools/profile_target.py book1 50 > $O/ncu_b1b.log 2>&1
# 2. book-2 final scene, wavefront mode: launch list + full set of k_extend / k_shade_all
run book2 8 > $O/plain_b2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file $O/launches_book2.csv python tools/profile_target.py book2 8 > $O/ncu_b2a.log 2>&1
run book2 8 > $O/plain_b2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'k_extend|k_shade_all' -s 20 -c 4 -o $O/prof_wave_book2 python tools/profile_target.py book2 8 > $O/ncu_b2b.log 2>&1
# 3. cornell smoke, wavefront
run smoke 20 > $O/plain_sm.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'k_extend|k_shade_all' -s 20 -c 2 -o $O/prof_wave_smoke python tools/profile_target.py smoke 20 > $O/ncu_sm.log 2>&1
# 4. 871k-triangle mesh room, fused
run mesh 2 > $O/plain_me.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'^k_mega(_r)?$' -c 1 -o $O/prof_mega_mesh python tools/profile_target.py mesh 2 > $O/ncu_me.log 2>&1
tail -n 1 $O/plain_b1.log $O/plain_b2.log $O/plain_sm.log $O/plain_me.log
ls -la $O/*.ncu-rep
# gpurun merges at most 64 MiB back: export the raw/source pages on the box and keep only the headline .ncu-rep
for r in prof_mega_book1 prof_wave_book2 prof_wave_smoke prof_mega_mesh; do
  [ -f $O/$r.ncu-rep ] || continue
  ncu -i $O/$r.ncu-rep --page raw --csv > $O/$r.raw.csv 2>/dev/null
  ncu -i $O/$r.ncu-rep --page source --csv > $O/$r.source.csv 2>/dev/null
  [ $r = prof_mega_book1 ] || rm -f $O/$r.ncu-rep
done
ls -la $O | head -40

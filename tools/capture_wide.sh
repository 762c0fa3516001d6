#!/bin/bash
# ncu evidence for the 4-wide fused kernels (run under gpurun, one GPU); every ncu command follows a plain run of the same command
set -u
O=gpurun_out
timeout 20 python tools/profile_target.py book1 50 > $O/plain_b1w.log 2>&1 && timeout 40 ncu --set full --clock-control none --import-source on -k regex:'^k_mega(_r)?$' -c 1 -o $O/prof_wide_book1 python tools/profile_target.py book1 50 > $O/ncu_b1w.log 2>&1
[ -f $O/prof_wide_book1.ncu-rep ] && ncu -i $O/prof_wide_book1.ncu-rep --page raw --csv > $O/prof_wide_book1.raw.csv 2>/dev/null
timeout 20 ncu --metrics gpu__time_duration.sum --clock-control none -c 50 --csv --log-file $O/launches_book1_wide.csv python tools/profile_target.py book1 50 > $O/ncu_b1wl.log 2>&1
timeout 20 python tools/profile_target.py mesh 2 > $O/plain_mew.log 2>&1 && timeout 45 ncu --set full --clock-control none --import-source on -k regex:'^k_mega(_r)?$' -c 1 -o $O/prof_wide_mesh python tools/profile_target.py mesh 2 > $O/ncu_mew.log 2>&1
[ -f $O/prof_wide_mesh.ncu-rep ] && ncu -i $O/prof_wide_mesh.ncu-rep --page raw --csv > $O/prof_wide_mesh.raw.csv 2>/dev/null
rm -f $O/prof_wide_mesh.ncu-rep
tail -n 1 $O/plain_b1w.log $O/plain_mew.log

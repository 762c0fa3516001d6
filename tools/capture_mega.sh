#!/bin/bash
set -u
O=gpurun_out
python tools/profile_target.py book1 50 > $O/plain_b1.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'^k_mega$' -c 1 -o $O/prof_mega_book1 python tools/profile_target.py book1 50 > $O/ncu_b1b.log 2>&1
python tools/profile_target.py mesh 2 > $O/plain_me.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'^k_mega$' -c 1 -o $O/prof_mega_mesh python tools/profile_target.py mesh 2 > $O/ncu_me.log 2>&1
ls -la $O/*.ncu-rep; tail -n 2 $O/ncu_b1b.log

#!/bin/bash
# ncu evidence for the pool kernel (run under gpurun, one GPU); every ncu command follows a plain run of the same command
set -u
O=gpurun_out
S=${1:-book1}; SPP=${2:-50}; TAG=${3:-pool_$S}
export RTB200_MODE=2
timeout 60 python tools/profile_target.py $S $SPP > $O/plain_$TAG.log 2>&1 && timeout 200 ncu --set full --clock-control none --import-source on -k regex:'^k_pool$' -c 1 -o $O/prof_$TAG python tools/profile_target.py $S $SPP > $O/ncu_$TAG.log 2>&1
[ -f $O/prof_$TAG.ncu-rep ] && ncu -i $O/prof_$TAG.ncu-rep --page raw --csv > $O/prof_$TAG.raw.csv 2>/dev/null
[ -f $O/prof_$TAG.ncu-rep ] && ncu -i $O/prof_$TAG.ncu-rep --page source --csv > $O/prof_$TAG.source.csv 2>/dev/null
rm -f $O/prof_$TAG.ncu-rep
tail -n 1 $O/plain_$TAG.log

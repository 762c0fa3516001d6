# round-2 late captures: the fused kernels after the signed-row walk and the multi-draw samplers (one gpurun call)
set -u
O=gpurun_out
python tools/ab_multi.py --scenes book1,mesh --reps 3 ${AB_LIBS:-_r0 _r1} 2>&1 | tee $O/r2_54_ab.log
run() { python tools/profile_target.py "$@"; }
run book1 50 > $O/plain_b1.log 2>&1 && timeout 200 ncu --set full --clock-control none --import-source on -k regex:'^k_mega(_r)?$' -c 1 -o $O/prof_r3_mega_book1 python tools/profile_target.py book1 50 > $O/ncu_b1.log 2>&1
if [ "${MESH:-0}" = 1 ]; then
run mesh 2 > $O/plain_me.log 2>&1 && timeout 200 ncu --set full --clock-control none --import-source on -k regex:'^k_mega(_r)?$' -c 1 -o $O/prof_r3_mega_mesh python tools/profile_target.py mesh 2 > $O/ncu_me.log 2>&1
fi
for r in prof_r3_mega_book1 prof_r3_mega_mesh; do
  [ -f $O/$r.ncu-rep ] || continue
  ncu -i $O/$r.ncu-rep --page raw --csv > $O/$r.raw.csv 2>/dev/null
  ncu -i $O/$r.ncu-rep --page source --csv > $O/$r.source.csv 2>/dev/null
  rm -f $O/$r.ncu-rep
done

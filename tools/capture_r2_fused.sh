set -u
O=gpurun_out
run() { python tools/profile_target.py "$@"; }
run book1 50 > $O/plain_b1.log 2>&1 && timeout 200 ncu --set full --clock-control none --import-source on -k regex:'^k_mega(_r)?$' -c 1 -o $O/prof_r2_mega_book1 python tools/profile_target.py book1 50 > $O/ncu_b1.log 2>&1
run mesh 2 > $O/plain_me.log 2>&1 && timeout 200 ncu --set full --clock-control none --import-source on -k regex:'^k_mega(_r)?$' -c 1 -o $O/prof_r2_mega_mesh python tools/profile_target.py mesh 2 > $O/ncu_me.log 2>&1
for r in prof_r2_mega_book1 prof_r2_mega_mesh; do
  [ -f $O/$r.ncu-rep ] || continue
  ncu -i $O/$r.ncu-rep --page raw --csv > $O/$r.raw.csv 2>/dev/null
  ncu -i $O/$r.ncu-rep --page source --csv > $O/$r.source.csv 2>/dev/null
  rm -f $O/$r.ncu-rep
done
python bench.py > $O/r2_bench_n1_b.json 2> $O/r2_bench_n1_b.err; tail -2 $O/r2_bench_n1_b.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/r2_bench_ref_b.json 2>&1

# last validation of the round: GPU suite, smoke, the default bench line and the reference arm (one gpurun call)
O=gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | tee $O/r2_gputests_7.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > $O/r2_bench_n1_k.json 2> $O/r2_bench_n1_k.err; tail -2 $O/r2_bench_n1_k.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/r2_bench_ref_k.json 2>&1
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2_bench_n1_k.json") if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["roofline"].get("frac_binding"), d["clocks"])
print({k:(v.get("paths_per_s"), v.get("render_wall_s")) for k,v in d.get("also",{}).items() if isinstance(v,dict) and "paths_per_s" in v})
PY
tail -c 600 $O/r2_bench_ref_k.json

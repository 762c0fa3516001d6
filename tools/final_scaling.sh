python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for n in 2 4 8; do python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2956$n bench.py --gpus $n --steps 5 --warmup 3 2> gpurun_out/r2_bench_n${n}_h.err | grep "^{" > gpurun_out/r2_bench_n${n}_h.json; done
python bench.py --no-extra --no-cpu-baseline 2>/dev/null | grep "^{" > gpurun_out/r2_bench_n1_h.json
python - <<PY
import json
for n in (1,2,4,8):
    d=json.load(open("gpurun_out/r2_bench_n%d_h.json"%n))
    print(n, d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d.get("also",{}).get("book2_final_strong",{}).get("render_wall_s"))
PY

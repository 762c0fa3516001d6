"""Attribute an ncu SASS source page (ncu -i X.ncu-rep --page source --csv) to CUDA source lines / functions.
usage: python tools/sass_lines.py source.csv k.sass [top]
k.sass = nvdisasm --print-line-info of the cubin (cuobjdump -xelf all librtb200.so).  The n-th instruction of the kernel
in the csv is matched with the instruction at the same offset in the disassembly."""
import csv, re, sys, collections, os, subprocess
src_csv, sass, top = sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40
rows = list(csv.reader(open(src_csv)))
kname = rows[0][1]
hdr = rows[1]; col = {n: i for i, n in enumerate(hdr)}
data = rows[2:]
base = int(data[0][0], 16)
# mangled-name pattern from the demangled template arguments
args = re.search(r"k_\w+<(.*?)>\(", kname)
kern = re.search(r"(k_\w+)<", kname).group(1)
vals = re.findall(r"\)(\d+)", args.group(1))
lines = open(sass).read().split("\n")
start = None
for i, l in enumerate(lines):
    if l.startswith("_ZN3rtb") and kern in l and l.endswith(":"):
        m = re.findall(r"L[bij](\d+)E", l.split("EEv")[0] + "E")
        if m == vals:
            start = i; break
assert start is not None, (kern, vals)
off2line = {}
cur = ("?", 0)
for l in lines[start + 1:]:
    if l.startswith("_ZN") and l.endswith(":"): break
    m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*);", l)
    if m: off2line[int(m.group(1), 16)] = (cur, m.group(2))
# function map from the source files
fn_of = {}
for f in ("rt_device.cuh", "kernels.cu"):
    p = os.path.join("ray_tracing_series_rust_b200", "csrc", "cuda", f)
    rev = os.environ.get("SRC_REV")  # the profile was taken from an older build: read the sources of that commit
    text = subprocess.run(["git", "show", f"{rev}:{p}"], capture_output=True, text=True, check=True).stdout if rev else open(p).read()
    name = "?"
    for n, l in enumerate(text.split("\n"), 1):
        m = re.match(r"^\s*(?:template\s*<[^>]*>\s*)?(?:__device__|__global__|static|RT_DEV|__forceinline__|inline|\s)+[\w:<>\*&\s]*?\b(\w+)\s*\([^;]*$", l)
        if m and ("__device__" in l or "__global__" in l or "RT_DEV" in l): name = m.group(1)
        fn_of[(f, n)] = name
agg_line = collections.Counter(); agg_fn = collections.Counter(); thr_fn = collections.Counter(); smp_fn = collections.Counter(); smp_line = collections.Counter()
tot_i = tot_t = tot_s = 0
for r in data:
    off = int(r[0], 16) - base
    (fl, _s) = off2line.get(off, (("?", 0), ""))
    ie = int(r[col["Instructions Executed"]] or 0); te = int(r[col["Thread Instructions Executed"]] or 0); sm = int(r[col["# Samples"]] or 0)
    fn = fn_of.get(fl, fl[0])
    agg_line[fl] += ie; agg_fn[fn] += ie; thr_fn[fn] += te; smp_fn[fn] += sm; smp_line[fl] += sm
    tot_i += ie; tot_t += te; tot_s += sm
print(f"kernel {kname}\ninstructions {tot_i:.3e}  thread-instr {tot_t:.3e}  avg threads {tot_t / max(tot_i, 1):.2f}  samples {tot_s}")
print("\nby function: inst%  samples%  threads/inst")
for fn, v in agg_fn.most_common(top):
    print(f"  {fn:28s} {100 * v / tot_i:6.2f} {100 * smp_fn[fn] / max(tot_s, 1):6.2f} {thr_fn[fn] / max(v, 1):6.2f}")
print("\nby line: samples%  inst%")
for fl, v in smp_line.most_common(top):
    print(f"  {fl[0]}:{fl[1]:<5d} {100 * v / max(tot_s, 1):6.2f} {100 * agg_line[fl] / tot_i:6.2f}  [{fn_of.get(fl, '?')}]")

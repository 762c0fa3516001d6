"""ncu launch list (--metrics gpu__time_duration.sum --csv --log-file X) -> per-kernel launches / total ms / share (profiles/*launches*.csv).
usage: python tools/summarise_launches.py launches.csv "comment line" > profiles/<tag>_launches_....csv"""
import csv, sys, collections, re
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot = collections.OrderedDict()
for r in rows[1:]:
    if len(r) <= vi or r[hdr.index("Metric Name")] != "gpu__time_duration.sum":
        continue
    v = float(r[vi].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[r[ui]]
    k = re.sub(r"\(.*$", "", r[ki]).replace("void ", "").replace("(unsigned int)", "").replace("(bool)", "").replace("(int)", "")
    n, t = tot.get(k, (0, 0.0))
    tot[k] = (n + 1, t + v)
s = sum(t for _, t in tot.values())
print("# " + (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]))
print("# ncu --metrics gpu__time_duration.sum --clock-control none; per-launch times are cold-cache and serialised: compare SHARES")
print("kernel,launches,total_ms,share_pct")
for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f'"{k}",{n},{t:.4f},{100 * t / s:.2f}')

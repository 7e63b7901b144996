"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel launches, time and share, split into
the setup phase (before the first PCG SpMM launch) and the PCG iterations.  Usage: summarize_launches.py in.csv[.gz] out.csv"""
import csv, gzip, io, re, sys
from collections import OrderedDict

src, dst = sys.argv[1], sys.argv[2]
raw = (gzip.open(src, "rt") if src.endswith(".gz") else open(src)).read()
rows = list(csv.DictReader(io.StringIO(raw[raw.index('"ID"'):])))
SPMM = re.compile(r"k_spmm")
phase, agg = "setup (mesh_set..rhs, before the first PCG iteration)", OrderedDict()
for r in rows:
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"^.*::", "", r["Kernel Name"].split("(")[0]).strip()
    if SPMM.search(name):
        phase = "PCG iterations"
    t = float(r["Metric Value"]) / 1e3
    k = (phase, name)
    n, s = agg.get(k, (0, 0.0))
    agg[k] = (n + 1, s + t)
tot = {}
for (ph, _), (_, s) in agg.items():
    tot[ph] = tot.get(ph, 0.0) + s
with open(dst, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["phase", "kernel", "launches", "time_us (ncu gpu__time_duration.sum, cold cache, serialised)", "share_of_phase"])
    for (ph, name), (n, s) in sorted(agg.items(), key=lambda kv: (kv[0][0] != "PCG iterations", -kv[1][1])):
        w.writerow([ph, name, n, "%.1f" % s, "%.4f" % (s / tot[ph])])
print({k: round(v, 1) for k, v in tot.items()})

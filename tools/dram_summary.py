#!/usr/bin/env python
"""Summarises an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv` capture of
tools/one_forward.py by kernel family for the LAST forward in the capture (it starts at the last-but-one stem kernel).
    python tools/dram_summary.py gpurun_out/dram_v3.csv [algorithmic GB per batch]"""
import collections, csv, re, sys
by = {}
with open(sys.argv[1]) as f:
    lines = [ln for ln in f if ln.startswith('"')]
for r in csv.DictReader(lines):
    d = by.setdefault(int(r["ID"]), {"name": r["Kernel Name"]})
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    if r["Metric Name"].startswith("dram"):
        d[r["Metric Name"]] = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    else:
        d["t"] = v / 1000 if u in ("ns", "nsecond") else (v if u.startswith("us") else v * 1000)
ids = sorted(by)
stems = [i for i in ids if "stem" in by[i]["name"]]
start = stems[-2] if len(stems) >= 2 else ids[0]
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for i in ids:
    if i < start:
        continue
    d = by[i]
    fam = re.sub(r"<.*", "", d["name"].split("(")[0]).replace("void ", "").replace("dyk::", "").strip() or "(torch)"
    a = agg[fam]
    a[0] += 1
    a[1] += d["t"]
    a[2] += d.get("dram__bytes_read.sum", 0)
    a[3] += d.get("dram__bytes_write.sum", 0)
tot = [0, 0.0, 0.0, 0.0]
print("one forward + NMS (launches %d..%d); times are cold-cache / serialised (ncu), traffic is what the metric names say" % (start, ids[-1]))
for fam, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"  {fam:26s} n={a[0]:4d}  {a[1] / 1000:7.3f} ms  read {a[2] / 1e9:7.3f} GB  write {a[3] / 1e9:7.3f} GB  ({(a[2] + a[3]) / a[1] / 1e3:7.1f} GB/s)")
    for k in range(4):
        tot[k] += a[k]
print(f"  {'TOTAL':26s} n={tot[0]:4d}  {tot[1] / 1000:7.3f} ms  read {tot[2] / 1e9:7.3f} GB  write {tot[3] / 1e9:7.3f} GB")
if len(sys.argv) > 2:
    alg = float(sys.argv[2])
    print(f"  algorithmic bytes per batch (SURVEY.md §8d): {alg:.2f} GB -> DRAM traffic / algorithmic = {(tot[2] + tot[3]) / 1e9 / alg:.2f}")

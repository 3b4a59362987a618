#!/usr/bin/env python
"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel family.
    python tools/launch_summary.py gpurun_out/launches.csv [first_launch] [last_launch]"""
import csv, re, sys
rows = []
with open(sys.argv[1]) as f:
    lines = [ln for ln in f if ln.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        us = v / 1000 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000)
        rows.append((int(r["ID"]), r["Kernel Name"], us))
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else 10 ** 9
rows = [r for r in rows if lo <= r[0] <= hi]
agg = {}
for _, name, us in rows:
    fam = re.sub(r"<.*", "", name.split("(")[0]).replace("void ", "").replace("dyk::", "").strip()
    a = agg.setdefault(fam, [0, 0.0])
    a[0] += 1
    a[1] += us
tot = sum(a[1] for a in agg.values())
print(f"{len(rows)} launches, {tot / 1000:.3f} ms of kernel time (cold-cache, serialised under ncu: compare shares)")
for fam, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"  {us / 1000:9.3f} ms {100 * us / tot:5.1f}%  n={n:5d}  {fam}")

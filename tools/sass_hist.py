#!/usr/bin/env python
"""SASS opcode histogram per kernel of libdyk_b200.so (cuobjdump -sass): the tensor-core / TMA / TMEM mnemonics that prove
which hardware path each kernel uses (UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA
load / store, UTCBAR = tcgen05.commit, SYNCS = mbarrier).    python tools/sass_hist.py > profiles/r02_sass_histogram.txt"""
import collections, re, subprocess, sys
from pathlib import Path
lib = Path(__file__).resolve().parent.parent / "double-yolo-kaist_b200" / "libdyk_b200.so"
KEY = ("UTCHMMA", "UTCQMMA", "UTCBAR", "UTCCP", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTMACCTL", "SYNCS", "HMMA", "MUFU",
       "LDG", "STG", "LDS", "STS", "BRX", "ATOM", "RED", "F2FP", "HFMA2", "FFMA", "FFMA2", "LDGSTS", "HMNMX2", "SHFL", "BAR", "UCGABAR", "ACQBULK", "LDSM")
proc = subprocess.Popen(["cuobjdump", "-sass", str(lib)], stdout=subprocess.PIPE, text=True)
hist, name = {}, None
for line in proc.stdout:
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name)
        hist[name] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", line)
    if m and name:
        hist[name][m.group(1)] += 1
        if m.group(1) in ("UTCHMMA", "UTCBAR", "UTMALDG", "LDTM") and m.group(2):
            hist[name][m.group(1) + m.group(2)] += 1
tot = collections.Counter()
print(f"# {lib.name}: {len(hist)} kernels; columns = total SASS instructions, then the hardware-path mnemonics present")
for n in sorted(hist):
    h = hist[n]
    total = sum(v for k, v in h.items() if "." not in k)
    keys = [k for k in h if k.split(".")[0] in KEY]
    for k in h:
        if "." not in k:
            tot[k] += h[k]
    main = {k: h[k] for k in sorted(keys) if "." not in k}
    variants = {k: h[k] for k in sorted(keys) if "." in k and ("2CTA" in k or "MULTICAST" in k)}
    print(f"{n}\n    {total} instr  " + " ".join(f"{k}={v}" for k, v in main.items()) + ("  | " + " ".join(f"{k}={v}" for k, v in variants.items()) if variants else ""))
print("\n# whole library")
print(" ".join(f"{k}={tot[k]}" for k in KEY if tot[k]))

#!/usr/bin/env python
"""Summarise an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line:
warp instructions executed, stall samples and the dominant stall reasons.
usage: ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass > x.csv; python scripts/ncu_lines.py x.csv [top]"""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
fname, hdr = None, None
agg = defaultdict(lambda: [0, 0, defaultdict(int), ""])
cur = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        ie, si = hdr.index("Instructions Executed"), hdr.index("# Samples")
        stalls = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    if r[0] != "":
        cur = (fname, int(r[0]))
        agg[cur][3] = r[1].strip()[:90]
    if cur is None or r[2] == "":
        continue
    try:
        agg[cur][0] += int(r[ie])
        agg[cur][1] += int(r[si])
        for i, h in stalls:
            if r[i] not in ("", "0"):
                agg[cur][2][h] += int(r[i])
    except ValueError:
        pass
tot_i = sum(v[0] for v in agg.values())
tot_s = sum(v[1] for v in agg.values())
print(f"total warp-instructions {tot_i}  samples {tot_s}")
bysmp = len(sys.argv) > 3 and sys.argv[3] == "samples"
for (f, ln), v in sorted(agg.items(), key=lambda kv: -kv[1][1 if bysmp else 0])[:top]:
    st = ",".join(f"{k[6:]}:{n}" for k, n in sorted(v[2].items(), key=lambda kv: -kv[1])[:3])
    print(f"{100*v[0]/tot_i:5.1f}% inst {100*v[1]/max(1,tot_s):5.1f}% smp  {f}:{ln:<4} {v[3]}   [{st}]")

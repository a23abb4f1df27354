#!/usr/bin/env python
"""Per-segment stall summary of one kernel from `ncu --page source --csv --print-source sass` output: segments end at
BAR / EXIT instructions (the hand-over points of the warp-specialised kernels).
usage: ncu -i X.ncu-rep --page source --csv --print-source sass > /tmp/src.csv; python tools/ncu_segments.py /tmp/src.csv [execs_per_unit]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
unit = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hdr, data = rows[1], rows[2:]
iS, iSrc, iX = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
st = [h for h in hdr if h.startswith("stall_") and "(Not" not in h]
idx = {h: hdr.index(h) for h in st}
print("total samples", sum(int(r[iS]) for r in data))
cur, n, ex, smp = {}, 0, 0, 0
for k, r in enumerate(data):
    for h in st:
        cur[h[6:]] = cur.get(h[6:], 0) + int(r[idx[h]])
    n += 1; ex += int(r[iX]); smp += int(r[iS])
    if 'BAR' in r[iSrc] or 'EXIT' in r[iSrc]:
        print(f"row {k:5d} {r[iSrc].strip()[:36]:36s} samples {smp:6d} sass {n:5d} exec/unit {ex / unit:8.1f}",
              {k2: v for k2, v in sorted(cur.items(), key=lambda kv: -kv[1]) if v > 50})
        cur, n, ex, smp = {}, 0, 0, 0

#!/usr/bin/env python
"""Per-source-line-range summary (stall samples, instructions, shared-memory wavefronts, opcode mix) of one kernel from
`ncu -i X.ncu-rep --page source --csv --print-source sass,cuda` output.
usage: python tools/ncu_sections.py src.csv file.cu "name:lo-hi,name:lo-hi,..."  (lines outside any range: by file name)"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
main = sys.argv[2]
ranges = []
for part in sys.argv[3].split(","):
    name, lohi = part.split(":"); lo, hi = lohi.split("-"); ranges.append((name, int(lo), int(hi)))
hdr = next(r for r in rows if r and r[0] == "Line No" and "# Samples" in r)
iS, iX, iW = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("L1 Wavefronts Shared")
cur = curfile = None
sec = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": curfile = r[1].split("/")[-1]; continue
    if r[0] in ("Function Name", "Line No"): continue
    if r[0] != "": cur = (curfile, int(r[0])); continue
    try: s, x = int(r[iS]), int(r[iX])
    except (ValueError, IndexError): continue
    try: wv = int(r[iW])
    except ValueError: wv = 0
    op = [o for o in r[3].split() if not o.startswith("@")]
    op = op[0] if op else ""
    op = ".".join(op.split(".")[:2]) if op.startswith(("LDS", "STS", "SHFL")) else op.split(".")[0]
    f, l = cur if cur else ("?", 0)
    name = f
    if f == main:
        name = next((n for n, lo, hi in ranges if lo <= l <= hi), f"{f}:other")
    b = sec.setdefault(name, [0, 0, 0, {}])
    b[0] += s; b[1] += x; b[2] += wv; b[3][op] = b[3].get(op, 0) + x
ts, tx, tw = (sum(b[i] for b in sec.values()) for i in range(3))
print(f"total samples {ts} warp-instructions {tx} shared wavefronts {tw}")
for name, b in sorted(sec.items(), key=lambda kv: -kv[1][0]):
    top = sorted(b[3].items(), key=lambda kv: -kv[1])[:9]
    print(f"{name:28s} samples {100 * b[0] / ts:5.1f}%  inst {100 * b[1] / tx:5.1f}%  wavefronts {100 * b[2] / max(tw, 1):5.1f}%  ", " ".join(f"{o}:{100 * x / tx:.1f}" for o, x in top))

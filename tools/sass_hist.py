#!/usr/bin/env python
"""Static opcode histogram of one kernel in an object file / library.  usage: python tools/sass_hist.py file.o kernel-substring"""
import collections, re, subprocess, sys
out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
cur, hist = None, collections.Counter()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m: cur = m.group(1); continue
    if cur and sys.argv[2] in cur:
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)(\.[A-Z0-9_.]+)?", line)
        if m:
            op, mod = m.group(1), m.group(2) or ""
            if op in ("LDS", "STS", "SHFL", "LDG", "STG"): op += "." + ".".join(mod.strip(".").split(".")[:1]) if mod else ""
            hist[op] += 1
tot = sum(hist.values())
print("total", tot, " ".join(f"{k}:{v}" for k, v in hist.most_common(24)))

#!/usr/bin/env python
"""Static SASS summary of libgeoac_b200.so, one row per kernel: registers, spill loads / stores (LDL / STL), TMA bulk copies
(UBLKCP), FP64 opcodes (DFMA / DMUL / DADD, fused share), address arithmetic (IMAD / IADD3 / LEA), global / shared loads and
stores, MUFU.  Runs in the build container (cuobjdump only, no GPU).  Usage: python scripts/sass_summary.py [lib.so] > profiles/rN_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "geoac_b200", "_lib", "libgeoac_b200.so")

res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
regs = {}
name = None
for line in res.splitlines():
    m = re.match(r"\s*Function (\S+):", line)
    if m:
        name = m.group(1)
        continue
    m = re.search(r"REG:(\d+).*?SHARED:(\d+).*?LOCAL:(\d+)", line)
    if m and name:
        regs[name] = (int(m.group(1)), int(m.group(2)), int(m.group(3)))

sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
demangle = lambda s: subprocess.run(["c++filt", s], capture_output=True, text=True).stdout.strip()
counts = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and cur:
        counts[cur][m.group(1)] += 1
        counts[cur]["_total"] += 1

cols = ["_total", "DFMA", "DMUL", "DADD", "MUFU", "IMAD", "IADD3", "LEA", "LDG", "STG", "LDS", "STS", "LDL", "STL", "UBLKCP", "SHFL", "BAR", "LDC"]
print(f"# {os.path.relpath(lib, ROOT)}  (cuobjdump -sass / -res-usage; static instruction counts, not dynamic)")
print("kernel | regs | local B | " + " | ".join(c.strip("_") for c in cols) + " | fused FP64 share")
for k, c in counts.items():
    short = demangle(k)
    short = re.sub(r"geoac::", "", short)
    short = re.sub(r"\(.*\)$", "", short)
    if not any(t in short for t in ("trace_kernel", "scout_kernel", "long_ray", "warp_ray")):
        continue
    r = regs.get(k, (0, 0, 0))
    fp = c["DFMA"] + c["DMUL"] + c["DADD"]
    print(f"{short} | {r[0]} | {r[2]} | " + " | ".join(str(c[x]) for x in cols) + f" | {c['DFMA'] / fp:.3f}" if fp else f"{short} | {r[0]}")

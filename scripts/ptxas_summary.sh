#!/bin/bash
# Print registers / spills / stack per kernel of libgeoac_b200 (nvcc -Xptxas -v), demangled and one line each.
cd "$(dirname "$0")/.."
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared -Xptxas -v "$@" \
     geoac_b200/csrc/capi.cu -o /tmp/ptxas_probe.so 2>&1 | python3 -c '
import sys, re, subprocess
name=None
for line in sys.stdin:
    m=re.search(r"Compiling entry function .(\S+). for", line)
    if m: name=subprocess.run(["c++filt", m.group(1)],capture_output=True,text=True).stdout.strip(); name=re.sub(r"\(.*","",name); continue
    m=re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m and name: stack=m.groups()
    m=re.search(r"Used (\d+) registers", line)
    if m and name: print(f"{name:75s} regs={m.group(1):>3s} stack={stack[0]} spill_st={stack[1]} spill_ld={stack[2]}"); name=None
'

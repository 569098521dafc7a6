"""Register / spill / shared-memory figures of every kernel from the -Xptxas -v logs the Makefile keeps
(lqr.jl_b200/csrc/*.ptxas.log).  Usage: python tools/ptxas_summary.py > profiles/r2_ptxas_resource_usage.txt"""
import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for log in sorted(glob.glob(os.path.join(ROOT, "lqr.jl_b200", "csrc", "*.ptxas.log"))):
    print("== " + os.path.basename(log))
    name = frame = None
    for line in open(log):
        m = re.search(r"Compiling entry function '([^']+)'", line)
        if m:
            name, frame = m.group(1), None
            continue
        if "bytes stack frame" in line:
            frame = line.strip()
            continue
        m = re.search(r"ptxas info\s+: (Used \d+ registers.*)", line)
        if m and name:
            print(f"{name}\t{frame or ''}\t{m.group(1).strip()}")
            name = None

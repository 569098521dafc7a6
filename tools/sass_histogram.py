"""SASS opcode histogram per kernel of liblqrb200.so (cuobjdump -sass): the evidence that the kernels are
Blackwell-native where the operation allows it — DMMA (FP64 tensor pipe; tcgen05 has no f64 kind), UBLKCP (1-D TMA bulk
copies), SYNCS (mbarrier), LDGSTS (cp.async), DFMA.  Usage: python tools/sass_histogram.py > profiles/r2_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "lqr.jl_b200", "liblqrb200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
kern, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern.replace("(anonymous namespace)::", ""))
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
KEYS = ["DMMA", "DFMA", "DMUL", "DADD", "MUFU", "UBLKCP", "SYNCS", "LDGSTS", "SHFL", "LDS", "STS", "LDG", "STG", "BAR"]
print("# SASS opcode counts per kernel (static), sm_100a, liblqrb200.so; total = all instructions")
print(f"{'kernel':110s} " + " ".join(f"{k:>6s}" for k in KEYS) + "  total")
tot = collections.Counter()
for k, c in hist.items():
    if sum(c.values()) < 50:
        continue
    print(f"{k[:110]:110s} " + " ".join(f"{c.get(x, 0):6d}" for x in KEYS) + f" {sum(c.values()):6d}")
    tot.update(c)
print(f"{'ALL KERNELS':110s} " + " ".join(f"{tot.get(x, 0):6d}" for x in KEYS) + f" {sum(tot.values()):6d}")
print("# no UTCMMA / tcgen05 instruction: there is no FP64 kind of tcgen05.mma; the FP64 tensor path is DMMA (mma.sync.m8n8k4.f64)")

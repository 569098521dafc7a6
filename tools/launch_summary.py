"""Per-kernel totals of an ncu launch list (--metrics gpu__time_duration.sum --csv).
usage: python tools/launch_summary.py gpurun_out/r2_bench_launches.csv > profiles/r2_bench_launches_summary.txt"""
import collections
import csv
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt, mx = collections.Counter(), collections.Counter(), collections.Counter()
for r in rows[1:]:
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0)
    k = r[ki]
    tot[k] += v
    cnt[k] += 1
    mx[k] = max(mx[k], v)
allv = sum(tot.values())
print("ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 python bench.py --steps 2 --warmup 3 --no-cpu-baseline")
print("(first 2,000 launches of the run = fp64 peak diagnostic, config 2 (incl. the e2e leg) and the start of config 1; cold-cache,")
print(" serialised: compare shares, not absolutes; torch / cutlass kernels = synthetic data generation, outside every timed region)")
for k, v in tot.most_common():
    print(f"{k[:110]:110s} n={cnt[k]:5d} total={v:12.1f} us share={100 * v / allv:5.1f}% max={mx[k]:10.1f} us")

#!/bin/bash
# ncu captures of the dominant kernel of each bench config (run AFTER the plain bench has exited 0).
#   usage: tools/gpu_profile.sh <tag> <config:kernel-regex[:probe-option]> ...   e.g.  r2 5aR:riccati_dmma c3:kkt_tpi
# with a probe option the workload is tools/perf_probe.py (replicated base batch) instead of bench.py
tag=$1; shift
for spec in "$@"; do
  IFS=: read cfg kern opt <<< "$spec"
  out=gpurun_out/${tag}_${cfg}_${kern}
  if [ -n "$opt" ]; then
    cmd="python tools/perf_probe.py --which $cfg --steps 2 --warmup 1 --opt $opt"
  else
    cmd="python bench.py --configs $cfg --steps 3 --warmup 3 --no-cpu-baseline"
  fi
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$kern -c 1 -f -o $out $cmd > $out.log 2>&1
  echo "== $spec rc=$? $(ls -la $out.ncu-rep 2>/dev/null | awk '{print $5}') bytes"
done

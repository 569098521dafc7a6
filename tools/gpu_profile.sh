#!/bin/bash
# ncu captures of the dominant kernel of each bench config (run AFTER the plain bench has exited 0).
#   usage: tools/gpu_profile.sh <tag> <config:kernel-regex> ...      e.g.  r2 5aR:riccati_dmma c3:kkt_tpi
tag=$1; shift
for spec in "$@"; do
  cfg=${spec%%:*}; kern=${spec#*:}
  out=gpurun_out/${tag}_${cfg}_${kern}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$kern -c 1 -f -o $out \
      python bench.py --configs $cfg --steps 3 --warmup 3 --no-cpu-baseline > $out.log 2>&1
  echo "== $spec rc=$? $(ls -la $out.ncu-rep 2>/dev/null | awk '{print $5}') bytes"
done

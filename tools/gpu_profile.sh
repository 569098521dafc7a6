#!/bin/bash
# ncu captures of the dominant kernel of each bench config (run AFTER the plain bench has exited 0); only the text
# summary of each capture (tools/ncu_summary.py) is kept under gpurun_out/ — the .ncu-rep files are 5-40 MB each.
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
  timeout 900 ncu --set full --clock-control none -k regex:$kern -c 1 -f -o /tmp/ncu_$$ $cmd > $out.log 2>&1
  rc=$?
  python tools/ncu_summary.py /tmp/ncu_$$.ncu-rep > ${out}_ncu_full.txt 2>&1
  echo "command: ncu --set full --clock-control none -k regex:$kern -c 1 $cmd" >> ${out}_ncu_full.txt
  rm -f /tmp/ncu_$$.ncu-rep; tail -c 300 $out.log > $out.tail; rm -f $out.log
  echo "== $spec rc=$rc"
done

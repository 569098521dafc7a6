#!/bin/bash
# compute-sanitizer over every kernel family (small batches); summaries go to gpurun_out/sanitizer_*.txt
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck synccheck; do
  for part in riccati kkt factor sqp; do
    out=gpurun_out/sanitizer_${tool}_${part}.txt
    timeout 900 $CS --tool $tool --print-limit 20 python tools/sanitize_driver.py $part > $out 2>&1
    echo "rc=$?" >> $out
    echo "== $tool $part: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|SANITIZE_DRIVER_OK|rc=' $out | tr '\n' ' ')"
  done
done

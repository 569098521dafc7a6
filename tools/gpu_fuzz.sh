#!/bin/bash
timeout 1200 python tools/fuzz_parity.py --cases ${1:-500} --seed ${2:-1} --budget-s 600 --out gpurun_out/r2_fuzz_seed${2:-1}.jsonl > gpurun_out/r2_fuzz_seed${2:-1}.log 2>&1
echo "fuzz rc=$?"; tail -n 40 gpurun_out/r2_fuzz_seed${2:-1}.log | cut -c 1-1500

#!/bin/bash
python -m pytest tests -m gpu -q > gpurun_out/r2_pytest6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest6.log; tail -25 gpurun_out/r2_pytest6.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench6.json 2> gpurun_out/r2_bench6.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r2_bench6.err

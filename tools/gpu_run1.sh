#!/bin/bash
# GPU run: full -m gpu test suite, the bench line and the conditioning stress grid
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest1.log; tail -5 gpurun_out/r2_pytest1.log
(time python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err); echo "bench rc=$?"; tail -c 1500 gpurun_out/r2_bench1.err
python tools/stress_scales.py > gpurun_out/r2_stress1.log 2>&1; tail -8 gpurun_out/r2_stress1.log

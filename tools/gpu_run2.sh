#!/bin/bash
# GPU run 2: full -m gpu suite (no -x), conditioning calibration
python -m pytest tests -m gpu -q > gpurun_out/r2_pytest2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest2.log; tail -30 gpurun_out/r2_pytest2.log
python tools/stress_scales.py > gpurun_out/r2_stress2.log 2>&1; tail -80 gpurun_out/r2_stress2.log

#!/bin/bash
# GPU run 3: full -m gpu suite, conditioning calibration (fixed tool), compute-sanitizer on every kernel family
python -m pytest tests -m gpu -q > gpurun_out/r2_pytest3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest3.log; tail -15 gpurun_out/r2_pytest3.log
python tools/stress_scales.py > gpurun_out/r2_stress3.log 2>&1; grep -c "above" gpurun_out/r2_stress3.log
bash tools/gpu_sanitize.sh

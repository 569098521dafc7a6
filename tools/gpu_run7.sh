#!/bin/bash
python -m pytest tests -m gpu -q > gpurun_out/r2_pytest7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest7.log; tail -30 gpurun_out/r2_pytest7.log

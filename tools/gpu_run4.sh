#!/bin/bash
python -m pytest tests -m gpu -q > gpurun_out/r2_pytest4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest4.log; tail -12 gpurun_out/r2_pytest4.log
bash tools/gpu_profile.sh r2a 5aR:riccati_dmma c4:dubins_sqp_step 5aK:kkt_hw2

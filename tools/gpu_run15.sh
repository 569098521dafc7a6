#!/bin/bash
python -m pytest tests -q -m gpu -x 2>&1 | tail -5
python tools/perf_probe.py --which 5bK,5bKp1,5bKp2,5bKp4,5aKp1 --steps 3 --warmup 1 2>&1 | tail -6
bash tools/gpu_run12.sh 4000 5 2>&1 | tail -4

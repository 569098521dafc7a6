#!/bin/bash
python -m pytest tests -q -m gpu -x 2>&1 | tail -5
python tools/perf_probe.py --which 5aR,c2 --steps 3 --warmup 1 2>&1 | tail -2
bash tools/gpu_run12.sh 4000 6 2>&1 | tail -4

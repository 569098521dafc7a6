#!/bin/bash
python -m pytest tests -q -m gpu 2>&1 | tail -4
bash tools/gpu_run12.sh 6000 13 2>&1 | tail -3
python tools/perf_probe.py --which 5aK,5bK --steps 3 --warmup 1 2>&1 | tail -2 | cut -c 1-220

#!/bin/bash
bash tools/gpu_run12.sh 5000 8 2>&1 | tail -3
python -m pytest tests -q -m gpu -x 2>&1 | tail -3

#!/bin/bash
python -m pytest tests/test_gpu_conditioning.py -q -m gpu -k "grid_stride or several_chunks" 2>&1 | tail -12

#!/bin/bash
python -m pytest tests -q -m gpu -x -k "on_the_cta_kernels" 2>&1 | tail -30

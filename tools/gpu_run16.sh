#!/bin/bash
python -m pytest tests -q -m gpu -k "test_cooperative_kernel or device_resident" 2>&1 | tail -4

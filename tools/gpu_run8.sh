#!/bin/bash
python -m pytest tests/test_gpu_kkt.py tests/test_gpu_determinism.py tests/test_gpu_conditioning.py -m gpu -q > gpurun_out/r2_pytest8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest8.log; tail -8 gpurun_out/r2_pytest8.log
echo "== 5aK kkt_variant=5 (kkt_wp, paired inversion + Hi record)"; python tools/perf_probe.py --which 5aK --steps 5 --opt kkt_variant=5
echo "== 5aK default (kkt_hw2)"; python tools/perf_probe.py --which 5aK --steps 5

#!/bin/bash
python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest5.log; tail -25 gpurun_out/r2_pytest5.log
echo "== 5aK default (kkt_wp)"; python tools/perf_probe.py --which 5aK --steps 5
echo "== 5aK kkt_variant=4 (kkt_hw2)"; python tools/perf_probe.py --which 5aK --steps 5 --opt kkt_variant=4
echo "== c4 default (prefetch ring)"; python tools/perf_probe.py --which c4 --steps 5
echo "== c4 sqp_prefetch=0"; python tools/perf_probe.py --which c4 --steps 5 --opt sqp_prefetch=0

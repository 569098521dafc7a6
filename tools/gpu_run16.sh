#!/bin/bash
python -m pytest tests -q -m gpu 2>&1 | tail -8
bash tools/gpu_run12.sh 5000 11 2>&1 | tail -3
for w in ric:10:3:101:16384 ric:14:7:101:8192 ric:20:6:101:4096 ric:30:8:101:2048 ric:7:2:101:32768; do
for pad in 1 0; do
python tools/perf_probe.py --which $w --steps 2 --warmup 1 --opt riccati_pad=$pad 2>&1 | tail -1 | python -c "import sys,json; r=json.loads(sys.stdin.read()); print(r['config'], r['kernel'], round(r['ms'],2), int(r['solves_per_s']))"
done; done

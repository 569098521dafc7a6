#!/bin/bash
python -m pytest tests -q -m gpu 2>&1 | tail -8
bash tools/gpu_run12.sh 5000 12 2>&1 | tail -3
for w in kkt:30:8:101:1024 kkt:20:8:101:2048 ric:20:10:101:4096 5bK 5bR; do
python tools/perf_probe.py --which $w --steps 2 --warmup 1 2>&1 | tail -1 | python -c "import sys,json; r=json.loads(sys.stdin.read()); print(r['config'], r['kernel'], round(r['ms'],2), int(r['solves_per_s']))"
done

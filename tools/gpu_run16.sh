#!/bin/bash
python -m pytest tests -q -m gpu 2>&1 | tail -6
for w in kkt:10:3:101:8192 kkt:14:7:101:4096 kkt:20:6:101:2048 kkt:30:8:101:1024 kkt:7:2:101:16384:1; do
for pad in 1 0; do
python tools/perf_probe.py --which $w --steps 2 --warmup 1 --opt kkt_pad=$pad 2>&1 | tail -1 | python -c "import sys,json; r=json.loads(sys.stdin.read()); print(r['config'], r['kernel'], round(r['ms'],2), int(r['solves_per_s']))"
done; done

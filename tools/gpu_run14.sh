#!/bin/bash
python tools/perf_probe.py --which 5aK --scale 0.125 --steps 2 --warmup 1 --opt kkt_variant=2 2>&1 | tail -2
python tools/perf_probe.py --which 5bK --scale 0.0625 --steps 2 --warmup 1 --opt kkt_variant=2 2>&1 | tail -2
python tools/perf_probe.py --which c3 --scale 0.125 --steps 2 --warmup 1 --opt kkt_variant=2 2>&1 | tail -2
python tools/perf_probe.py --which 5aR,5bR,c2 --scale 0.125 --steps 2 --warmup 1 --opt riccati_variant=2 2>&1 | tail -4
ncu --set full --clock-control none -k regex:kkt_coop -c 1 -o gpurun_out/coop5aK python tools/perf_probe.py --which 5aK --scale 0.03125 --steps 1 --warmup 0 --opt kkt_variant=2 > /dev/null 2>&1
ncu -i gpurun_out/coop5aK.ncu-rep --page details > gpurun_out/r2_kkt_coop_12_4_ncu_full.txt 2>&1; rm -f gpurun_out/coop5aK.ncu-rep
grep -E "Duration|Registers Per|Theoretical Occ|Achieved Occ|Executed Ipc|Issue Slots Busy|DRAM Throughput|Shared Memory Config|Dynamic Shared|Block Limit|Waves" gpurun_out/r2_kkt_coop_12_4_ncu_full.txt | head -20

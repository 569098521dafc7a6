#!/bin/bash
N=$1
if [ "$N" = "1" ]; then
  python tools/h2d_probe.py > gpurun_out/r2_h2d_probe_${N}gpu.json 2> gpurun_out/r2_h2d_probe_${N}gpu.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 tools/h2d_probe.py > gpurun_out/r2_h2d_probe_${N}gpu.json 2> gpurun_out/r2_h2d_probe_${N}gpu.err
fi
echo "probe ${N}gpu rc=$?"; cat gpurun_out/r2_h2d_probe_${N}gpu.json; tail -c 1500 gpurun_out/r2_h2d_probe_${N}gpu.err; nproc; lscpu | grep -i "numa\|model name\|socket"

#!/bin/bash
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err; echo "bench ${N}gpu rc=$?"; tail -c 300 gpurun_out/r2_bench_${N}gpu.err; free -g | head -2

#!/bin/bash
# 2-GPU check of the bench contract (the driver launches it the same way)
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err; echo "bench 2gpu rc=$?"; tail -c 500 gpurun_out/r2_bench_2gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2_bench_2gpu_ref.json 2> gpurun_out/r2_bench_2gpu_ref.err; echo "ref 2gpu rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2

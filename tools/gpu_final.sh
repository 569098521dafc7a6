#!/bin/bash
# final measurement run: both bench arms, ncu launch list, ncu full captures (text summaries only)
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; echo "reference arm rc=$?"
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "bench rc=$?"; tail -c 400 gpurun_out/r2_bench_final.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /tmp/under_ncu.log 2>&1; echo "launch list rc=$?"
bash tools/gpu_profile.sh r2 c2:riccati_tpi c3:kkt_tpi c4:dubins_sqp_step 5aR:riccati_dmma 5aK:kkt_hw2 5aK:kkt_hinv \
    5bR:riccati_cta 5bK:kkt_cta_kernel 5bK:kkt_cta_prep 5aK:kkt_wp:kkt_variant=5
du -sh gpurun_out

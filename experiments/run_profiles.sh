#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:filter_group_agg -s 2 -c 1 -f -o gpurun_out/prof_agg_r01b python bench.py --no-e2e --no-cpu --no-knn --steps 1 --warmup 1 > gpurun_out/ncu_agg_r01b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:knn_gemm -s 5 -c 1 -f -o gpurun_out/prof_knn_r01b python experiments/knn_bench.py --n 10000000 --reps 1 > gpurun_out/ncu_knn_r01b.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/bench_launches_r01b.csv python bench.py --no-e2e --no-cpu --steps 2 --warmup 1 > gpurun_out/ncu_bench_r01b.log 2>&1
ls -la gpurun_out/*r01b*

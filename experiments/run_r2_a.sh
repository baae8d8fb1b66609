#!/bin/bash
# round 2, GPU call A (1 GPU): new agg kernels: parity, knob sweep, microbench, bench line with verification
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,persistence_mode,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 600 python -m pytest tests/test_agg_gpu.py tests/test_peer_exchange_procs_gpu.py tests/test_filter_gpu.py -x -q -m gpu > gpurun_out/r2a_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2a_tests.log
tail -15 gpurun_out/r2a_tests.log
timeout 300 python experiments/agg_sweep.py 1000000000 1000000 3 > gpurun_out/r2a_sweep.txt 2>&1
cat gpurun_out/r2a_sweep.txt
timeout 200 ./experiments/atomics_bench > gpurun_out/r2a_atomics.txt 2>&1
tail -12 gpurun_out/r2a_atomics.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-knn --no-cpu --e2e-rows 200000000 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
tail -c 3000 gpurun_out/r2a_bench.json; tail -5 gpurun_out/r2a_bench.err

#!/bin/bash
# round 2, GPU call B (1 GPU): whole GPU test suite, per-operator numbers, the bench line with verification
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 --maxfail 25 -p no:cacheprovider > gpurun_out/r2b_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2b_tests.log
tail -40 gpurun_out/r2b_tests.log
timeout 600 python experiments/bench_ops.py --reps 2 > gpurun_out/r2b_ops.jsonl 2> gpurun_out/r2b_ops.err
cat gpurun_out/r2b_ops.jsonl | cut -c1-900; tail -5 gpurun_out/r2b_ops.err
timeout 600 python bench.py --steps 10 --warmup 3 --small-block-rows 65536 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
tail -c 4000 gpurun_out/r2b_bench.json; tail -5 gpurun_out/r2b_bench.err

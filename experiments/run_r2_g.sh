#!/bin/bash
# round 2, GPU call G (1 GPU): kNN rework (tests + bench), join default, agg kernel capture
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 --maxfail 25 -p no:cacheprovider > gpurun_out/r2g_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2g_tests.log
tail -30 gpurun_out/r2g_tests.log
timeout 600 python experiments/bench_ops.py --ops join --reps 2 2>/dev/null | cut -c1-520
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
python - <<'P'
import json
d = json.loads(open("gpurun_out/r2g_bench.json").read().strip().splitlines()[-1])
print("agg ms/step", d["ms_per_step"], "frac", d["roofline"]["frac"], "verify", d["verify"]["ok"], "e2e", d["e2e"]["value"], "small", d["e2e"].get("small_blocks"))
k = d["knn"]; print("knn qps", k["value"], "ms", k["ms_per_batch"], "gemm", k["roofline"]["kernel_ms"], "frac", k["roofline"]["frac"], "launches", k["gpu_launches_per_batch"], "cert", k["certified_queries"], "exact", k["exact_fallback_queries"], "e2e", k["e2e"]["value"])
P
tail -3 gpurun_out/r2g_bench.err
DBX_KNN_SHARED_LIST=1 timeout 300 python experiments/knn_bench.py --n 10000000 --reps 3 2>&1 | tail -4
timeout 300 python experiments/knn_bench.py --n 10000000 --reps 3 2>&1 | tail -4
timeout 300 python experiments/knn_bench.py --n 1250000 --reps 5 2>&1 | tail -4

#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/agg_variants.log
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"])'
for cfg in "3 3" "3 6" "3 12" "2 2" "2 8" "4 8"; do
  set -- $cfg
  echo "== minb=$1 grid_per_sm=$2" >> gpurun_out/agg_variants.log
  DBX_AGG_MINB=$1 DBX_AGG_GRID=$2 timeout 200 python bench.py --no-e2e --no-cpu --no-knn --steps 5 --warmup 3 2>&1 | tail -1 | python -c "$P" >> gpurun_out/agg_variants.log 2>&1
done
cat gpurun_out/agg_variants.log

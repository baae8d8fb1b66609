#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/agg2.log
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["config"]["groups_out"])'
for keys in 1000000 500000 250000 100000 2000000; do
  echo "== keys=$keys default" >> gpurun_out/agg2.log
  timeout 300 python bench.py --no-e2e --no-cpu --steps 5 --warmup 3 --keys $keys 2>&1 | tail -1 | python -c "$P" >> gpurun_out/agg2.log 2>&1
done
echo "== keys=1e6 RED evict_last hint" >> gpurun_out/agg2.log
DBX_AGG_DEBUG=8 timeout 300 python bench.py --no-e2e --no-cpu --steps 5 --warmup 3 2>&1 | tail -1 | python -c "$P" >> gpurun_out/agg2.log 2>&1
echo "== keys=1e6 no table updates (debug 1)" >> gpurun_out/agg2.log
DBX_AGG_DEBUG=1 timeout 300 python bench.py --no-e2e --no-cpu --steps 5 --warmup 3 2>&1 | tail -1 | python -c "$P" >> gpurun_out/agg2.log 2>&1
cat gpurun_out/agg2.log

#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/agg_bulk.log
echo "== tests (bulk default)" >> gpurun_out/agg_bulk.log
timeout 600 python -m pytest tests/test_agg_gpu.py -x -q -m gpu 2>&1 | tail -5 >> gpurun_out/agg_bulk.log
echo "== tests (no bulk)" >> gpurun_out/agg_bulk.log
DBX_AGG_NO_BULK=1 timeout 600 python -m pytest tests/test_agg_gpu.py -x -q -m gpu 2>&1 | tail -3 >> gpurun_out/agg_bulk.log
for lanes in FFFFFFFF 77777777 55555555 11111111; do
  echo "== bench lanes=$lanes" >> gpurun_out/agg_bulk.log
  DBX_AGG_BULK_LANES=$lanes timeout 300 python bench.py --no-e2e --no-cpu --steps 5 --warmup 3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['value'])" >> gpurun_out/agg_bulk.log 2>&1
done
echo "== bench no bulk" >> gpurun_out/agg_bulk.log
DBX_AGG_NO_BULK=1 timeout 300 python bench.py --no-e2e --no-cpu --steps 5 --warmup 3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['value'])" >> gpurun_out/agg_bulk.log 2>&1
cat gpurun_out/agg_bulk.log

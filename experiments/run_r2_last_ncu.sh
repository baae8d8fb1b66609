#!/bin/bash
# ncu captures of the two kernels added late in round 2: the generated expression kernel and pass 1 of the two-pass aggregation
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_tag_requests.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active
timeout 200 ncu --metrics $M --clock-control none -k regex:dbx_jit_eval -s 1 -c 1 --csv --log-file gpurun_out/r2z_ncu_eval.csv python experiments/bench_ops.py --ops eval --reps 1 > /dev/null 2>&1
cat > /tmp/big_agg.py <<'P'
import os, sys
sys.path.insert(0, os.getcwd())
import torch
from databend_b200 import abi, build, lib, expr as E
from databend_b200.block import Column, DataBlock
from databend_b200.transforms import AggregatorParams, TransformFinalAggregate, TransformPartialAggregate
build.build(); L = lib.load(); lib.require_device()
rows, nk = 1 << 28, 4_000_000
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(3)
k = torch.randint(0, nk, (rows,), dtype=torch.int64, device=dev, generator=g)
v = torch.randint(0, 1 << 40, (rows,), dtype=torch.int64, device=dev, generator=g)
x = torch.randint(0, 1 << 20, (rows,), dtype=torch.int64, device=dev, generator=g).to(torch.float64)
blk = DataBlock([Column.device(abi.I64, rows, k.data_ptr()), Column.device(abi.I64, rows, v.data_ptr()), Column.device(abi.F64, rows, x.data_ptr())], rows)
p = AggregatorParams([0], [("sum", 1), ("count", 1), ("avg", 2)], expected_groups=nk)
part = TransformPartialAggregate(p, [abi.I64, abi.I64, abi.F64], E.eq(E.col(1) % E.lit(3), E.lit(0)))
for _ in range(2):
    part.reset(); part.transform(blk); part.on_finish()
print(part.kernel_variant(), part.last_kernel_ms())
P
timeout 300 ncu --metrics $M --clock-control none -k regex:'filter_partition_kernel|filter_group_agg_kernel' -s 9 -c 9 --csv --log-file gpurun_out/r2z_ncu_twopass.csv python /tmp/big_agg.py > gpurun_out/r2z_ncu_twopass.log 2>&1
tail -2 gpurun_out/r2z_ncu_twopass.log
python - <<'P'
import csv
for f in ("gpurun_out/r2z_ncu_eval.csv", "gpurun_out/r2z_ncu_twopass.csv"):
    rows = list(csv.reader(open(f)))
    hdr = None
    for r in rows:
        if len(r) > 5 and r[0] == "ID": hdr = r; continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            print(d["ID"], d["Kernel Name"][:44], d["Metric Name"], d["Metric Value"], d["Metric Unit"])
P

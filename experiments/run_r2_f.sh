#!/bin/bash
# round 2, GPU call F (1 GPU): tests after the kNN / sort changes, per-operator numbers, join ablations, the bench line, agg kernel capture
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 --maxfail 25 -p no:cacheprovider > gpurun_out/r2f_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2f_tests.log
tail -30 gpurun_out/r2f_tests.log
timeout 600 python experiments/bench_ops.py --reps 2 > gpurun_out/r2f_ops.jsonl 2> gpurun_out/r2f_ops.err
cut -c1-600 gpurun_out/r2f_ops.jsonl; tail -5 gpurun_out/r2f_ops.err
echo "== join without radix regions"; DBX_JOIN_REGION_BYTES=0 timeout 300 python experiments/bench_ops.py --ops join --reps 2 2>/dev/null | cut -c1-520
echo "== join, 64 MB regions"; DBX_JOIN_REGION_BYTES=67108864 timeout 300 python experiments/bench_ops.py --ops join --reps 2 2>/dev/null | cut -c1-520
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
python - <<'P'
import json
d = json.loads(open("gpurun_out/r2f_bench.json").read().strip().splitlines()[-1])
print("agg ms/step", d["ms_per_step"], "frac", d["roofline"]["frac"], "verify", d["verify"]["ok"], "e2e", d["e2e"]["value"], "small", d["e2e"].get("small_blocks"))
k = d["knn"]; print("knn qps", k["value"], "ms", k["ms_per_batch"], "gemm", k["roofline"]["kernel_ms"], "frac", k["roofline"]["frac"], "launches", k["gpu_launches_per_batch"], "cert", k["certified_queries"], "e2e", k["e2e"]["value"])
P
tail -3 gpurun_out/r2f_bench.err
NCU="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
timeout 300 $NCU --log-file gpurun_out/r2f_sort_launches.csv python experiments/bench_ops.py --ops sort --sort-rows 100000000 --reps 1 > gpurun_out/r2f_sort.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'filter_group_agg_kernel<.int.3, .bool.1' -s 2 -c 1 -f -o gpurun_out/r2f_prof_agg python bench.py --no-e2e --no-cpu --no-knn --no-verify --steps 1 --warmup 1 > gpurun_out/r2f_ncu_agg.log 2>&1
timeout 300 ncu --metrics lts__t_sectors_op_red.sum,lts__t_sectors_op_atom.sum,lts__t_requests_srcunit_tex_op_red.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum,l1tex__t_set_accesses_pipe_lsu_mem_global_op_red.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:'filter_group_agg_kernel<.int.3, .bool.1' -s 2 -c 1 --csv --log-file gpurun_out/r2f_agg_red_counters.csv python bench.py --no-e2e --no-cpu --no-knn --no-verify --steps 1 --warmup 1 > gpurun_out/r2f_ncu_agg2.log 2>&1
python - <<'P'
import csv, collections
rows = list(csv.reader(open("gpurun_out/r2f_sort_launches.csv")))
hdr = None
agg = collections.OrderedDict()
for r in rows:
    if len(r) > 5 and r[0] == "ID": hdr = r; continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        try: v = float(d["Metric Value"].replace(",", ""))
        except: continue
        k = d["Kernel Name"][:70]
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:8]:
    print(f"{k:70s} n={c:5d} total={t/1e6:9.3f} ms")
P
tail -2 gpurun_out/r2f_ncu_agg.log; cut -d, -f5,9,13,15 gpurun_out/r2f_agg_red_counters.csv | tail -14

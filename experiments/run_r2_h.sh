#!/bin/bash
# round 2, GPU call H (1 GPU): expression evaluator tests, aggregate tests with the run-time specialised kernels,
# bench line with and without specialisation, ncu capture + RED counters of the specialised kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_eval_gpu.py tests/test_agg_gpu.py tests/test_abi.py -q -m "gpu or not gpu" --timeout 600 --maxfail 10 -p no:cacheprovider > gpurun_out/r2h_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2h_tests.log
tail -25 gpurun_out/r2h_tests.log
for jit in 1 0; do
  DBX_AGG_JIT=$jit timeout 600 python bench.py --steps 10 --warmup 3 --no-knn > gpurun_out/r2h_bench_jit$jit.json 2> gpurun_out/r2h_bench_jit$jit.err
  python - <<P
import json
d = json.loads(open("gpurun_out/r2h_bench_jit$jit.json").read().strip().splitlines()[-1])
r = d["roofline"]
print("jit=$jit ms/step", d["ms_per_step"], "kernel_ms", r["kernel_ms"], "frac", r["frac"], r.get("kernel_variant"), "verify", d["verify"]["ok"] if d.get("verify") else None, "e2e", d["e2e"]["value"], d["e2e"].get("small_blocks"))
P
  tail -3 gpurun_out/r2h_bench_jit$jit.err
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'dbx_jit_agg_fast' -s 2 -c 1 -f -o gpurun_out/r2h_prof_agg_jit python bench.py --no-e2e --no-cpu --no-knn --no-verify --steps 1 --warmup 1 > gpurun_out/r2h_ncu_agg.log 2>&1
timeout 300 ncu --metrics lts__t_sectors_op_red.sum,lts__t_requests_srcunit_tex_op_red.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum,l1tex__t_set_accesses_pipe_lsu_mem_global_op_red.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:'dbx_jit_agg_fast' -s 2 -c 1 --csv --log-file gpurun_out/r2h_agg_jit_counters.csv python bench.py --no-e2e --no-cpu --no-knn --no-verify --steps 1 --warmup 1 > gpurun_out/r2h_ncu_agg2.log 2>&1
tail -2 gpurun_out/r2h_ncu_agg.log; cut -d, -f5,9,13,15 gpurun_out/r2h_agg_jit_counters.csv | tail -15

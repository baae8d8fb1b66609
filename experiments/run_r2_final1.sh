#!/bin/bash
# round 2, final single-GPU call: whole GPU suite, the bench line (both arms), per-operator numbers, key-count / skew
# variants, the launch list of the bench under ncu, a kNN GEMM capture, compute-sanitizer over the new code paths
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 --maxfail 20 -p no:cacheprovider > gpurun_out/r2z_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2z_tests.log; tail -6 gpurun_out/r2z_tests.log | cut -c1-200
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2z_bench_reference.json 2> gpurun_out/r2z_bench_reference.err
cut -c1-600 gpurun_out/r2z_bench_reference.json
timeout 900 python bench.py > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err
python - <<'P'
import json
d = json.loads(open("gpurun_out/r2z_bench.json").read().strip().splitlines()[-1])
r = d["roofline"]
print("agg ms/step", d["ms_per_step"], "kernel", r["kernel_ms"], "frac", r["frac"], r.get("kernel_variant"), "traffic", r.get("traffic"), "verify", d["verify"]["ok"], "clocks", d.get("clocks"))
print("e2e", d["e2e"]["value"], d["e2e"].get("small_blocks"), "cpu", d.get("cpu_baseline"))
k = d["knn"]; print("knn qps", k["value"], "ms", k["ms_per_batch"], "gemm", k["roofline"]["kernel_ms"], "frac", k["roofline"]["frac"], "e2e", k["e2e"]["value"])
P
tail -3 gpurun_out/r2z_bench.err
timeout 900 python experiments/bench_ops.py --reps 2 > gpurun_out/r2z_ops.jsonl 2> gpurun_out/r2z_ops.err
cut -c1-420 gpurun_out/r2z_ops.jsonl; tail -3 gpurun_out/r2z_ops.err
timeout 600 python experiments/agg_variants.py > gpurun_out/r2z_agg_variants.jsonl 2> gpurun_out/r2z_agg_variants.err
cat gpurun_out/r2z_agg_variants.jsonl; tail -3 gpurun_out/r2z_agg_variants.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2z_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-verify > gpurun_out/r2z_bench_under_ncu.log 2>&1
python - <<'P'
import csv, collections
rows = list(csv.reader(open("gpurun_out/r2z_bench_launches.csv")))
hdr = None
agg = collections.OrderedDict()
for r in rows:
    if len(r) > 5 and r[0] == "ID": hdr = r; continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        try: v = float(d["Metric Value"].replace(",", ""))
        except Exception: continue
        k = d["Kernel Name"][:60]
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(t for _, t in agg.values())
for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:12]:
    print(f"{k:60s} n={c:5d} total={t/1e6:9.3f} ms share={t/tot*100:5.1f}%")
P
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'knn_gemm_filter_kernel' -s 3 -c 1 -f -o gpurun_out/r2z_prof_knn python experiments/knn_bench.py --n 10000000 --reps 1 > gpurun_out/r2z_ncu_knn.log 2>&1
tail -2 gpurun_out/r2z_ncu_knn.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_eval_gpu.py tests/test_agg_spill_gpu.py tests/test_block_kernels_gpu.py "tests/test_agg_gpu.py::test_wide_128_bit_group_keys" "tests/test_agg_gpu.py::test_small_pinned_host_blocks_gathered_by_the_device" "tests/test_agg_gpu.py::test_specialised_kernels_serve_grouped_plans" -q -m gpu -x --timeout 900 -p no:cacheprovider -k "not all_type_pairs" > gpurun_out/r2z_sanitizer.log 2>&1
echo "sanitizer rc=$?" >> gpurun_out/r2z_sanitizer.log; tail -8 gpurun_out/r2z_sanitizer.log | cut -c1-200

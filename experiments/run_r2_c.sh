#!/bin/bash
# round 2, multi-GPU call: N ranks on one box.  usage: run_r2_c.sh N
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2y_topo_n$N.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
# (1) the driver's own command shape: full line (agg + e2e + knn), pipelined exchange
timeout 900 $TR --master-port 29541 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2y_bench_n$N.json 2> gpurun_out/r2y_bench_n$N.err
tail -c 3500 gpurun_out/r2y_bench_n$N.json; tail -3 gpurun_out/r2y_bench_n$N.err
# (2) agg leg only, without the cross-operator pipelining, with nvidia-smi polling every GPU in the background (as a driver would)
( nvidia-smi --query-gpu=index,clocks.sm,power.draw --format=csv,noheader -lms 200 > gpurun_out/r2y_smi_n$N.csv 2>&1 & echo $! > /tmp/smi.pid )
DBX_BENCH_PIPELINE=0 NCCL_DEBUG=INFO NCCL_DEBUG_FILE=gpurun_out/r2y_nccl_n$N.%p.log timeout 600 $TR --master-port 29542 bench.py --gpus $N --steps 20 --warmup 5 --no-e2e --no-knn --no-cpu > gpurun_out/r2y_bench_nopipe_n$N.json 2> gpurun_out/r2y_bench_nopipe_n$N.err
kill $(cat /tmp/smi.pid) 2>/dev/null
python - <<P
import json
for f in ["gpurun_out/r2y_bench_n$N.json", "gpurun_out/r2y_bench_nopipe_n$N.json"]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "ms/step", d["ms_per_step"], "value", d["value"], "kernel", d["roofline"]["kernel_ms"], "phases", d.get("phases"), "verify", (d.get("verify") or {}).get("ok"), "per_step", d["config"].get("per_step_wall_ms"))
    except Exception as e:
        print(f, "unreadable:", e)
P
tail -3 gpurun_out/r2y_bench_nopipe_n$N.err
rm -f gpurun_out/r2y_nccl_n$N.*.log
# (3) join (fused peer shuffle) and top-k (device merge) on N ranks
timeout 900 $TR --master-port 29543 experiments/bench_ops.py --ops join,topk --reps 2 > gpurun_out/r2y_ops_n$N.jsonl 2> gpurun_out/r2y_ops_n$N.err
cut -c1-1200 gpurun_out/r2y_ops_n$N.jsonl; tail -5 gpurun_out/r2y_ops_n$N.err

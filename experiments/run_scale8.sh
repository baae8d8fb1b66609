#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
bash experiments/run_scale.sh $N
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 experiments/bench_ops.py --ops join,topk --reps 2 > gpurun_out/ops_n${N}.jsonl 2> gpurun_out/ops_n${N}.err
python - <<P
import json
for l in open("gpurun_out/ops_n${N}.jsonl"):
    d=json.loads(l)
    print(d["op"], "N=", d["n_gpus"], "rows/s", d["rows_per_s"], "total_ms", d["total_ms"], {k:d[k] for k in d if k.endswith("_ms")})
P
tail -3 gpurun_out/ops_n${N}.err

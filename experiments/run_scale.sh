#!/bin/bash
# strong-scaling run of the agg leg at N GPUs (peer-memory exchange), plus the full line at N
N=${1:-8}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 5 --no-e2e --no-cpu --no-knn > gpurun_out/scale_n${N}_agg.json 2> gpurun_out/scale_n${N}_agg.err
python - <<P
import json
d=json.loads(open("gpurun_out/scale_n${N}_agg.json").read().strip().splitlines()[-1])
print("N=${N} agg ms/step", d["ms_per_step"], "wall", d["wall_ms_per_step"], "kernel", d["roofline"]["kernel_ms"], "value", d["value"], "launches", d["gpu_launches"])
P
tail -2 gpurun_out/scale_n${N}_agg.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu > gpurun_out/scale_n${N}_full.json 2> gpurun_out/scale_n${N}_full.err
python - <<P
import json
d=json.loads(open("gpurun_out/scale_n${N}_full.json").read().strip().splitlines()[-1])
print("N=${N} full: agg ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"] if d["e2e"] else None, "knn qps", d["knn"]["value"], "ms", d["knn"]["ms_per_batch"], "gemm ms", d["knn"]["roofline"]["kernel_ms"], "frac", d["knn"]["roofline"]["frac"])
P
tail -2 gpurun_out/scale_n${N}_full.err

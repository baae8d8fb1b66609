#!/bin/bash
# usage: experiments/run_n2.sh N   — torchrun bench on N GPUs, output in gpurun_out/
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-e2e > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "rc=$?"
tail -20 gpurun_out/bench_n$N.err
cat gpurun_out/bench_n$N.json

#!/bin/bash
# N-GPU bench on one box: peer-memory exchange vs the NCCL all-to-all path
N=${1:-2}
mkdir -p gpurun_out
for mode in peer nccl; do
  DBX_EXCHANGE=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-e2e --no-cpu --no-knn > gpurun_out/bench_n${N}_$mode.json 2> gpurun_out/bench_n${N}_$mode.err
  tail -c 1500 gpurun_out/bench_n${N}_$mode.json; tail -3 gpurun_out/bench_n${N}_$mode.err
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n${N}_full.json 2> gpurun_out/bench_n${N}_full.err
tail -c 2500 gpurun_out/bench_n${N}_full.json; tail -3 gpurun_out/bench_n${N}_full.err

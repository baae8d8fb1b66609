#!/bin/bash
# end-of-round validation: GPU test suite, smoke(), a short 2-GPU bench (peer exchange path)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -3 > gpurun_out/final_tests.log; cat gpurun_out/final_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
N=${1:-2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 --no-e2e --no-cpu --no-knn > gpurun_out/final_n${N}.json 2> gpurun_out/final_n${N}.err
python - <<P
import json
d=json.loads(open("gpurun_out/final_n${N}.json").read().strip().splitlines()[-1])
print("N=${N}", d["ms_per_step"], d["value"], d["config"]["parallelism"], d["roofline"]["kernel_ms"])
P
grep -v "OMP_NUM\|\*\*\*" gpurun_out/final_n${N}.err | tail -3

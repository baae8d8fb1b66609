#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/knn_tests.log gpurun_out/knn_bench.log
for c in 1 2 4; do
  echo "== tests cluster=$c" >> gpurun_out/knn_tests.log
  DBX_KNN_CLUSTER=$c timeout 300 python -m pytest tests/test_knn_gpu.py -x -q -m gpu -k "random_768 or ragged or tensor_core or duplicates" 2>&1 | tail -15 >> gpurun_out/knn_tests.log
done
for c in 1 2 4 8; do
  echo "== bench cluster=$c n=10M" >> gpurun_out/knn_bench.log
  DBX_KNN_CLUSTER=$c timeout 300 python experiments/knn_bench.py --n 10000000 --reps 3 2>&1 | tail -3 >> gpurun_out/knn_bench.log
done
DBX_KNN_CLUSTER=2 timeout 300 python experiments/knn_bench.py --n 10000000 --reps 2 --fn l2_distance 2>&1 | tail -2 >> gpurun_out/knn_bench.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:knn_gemm -s 4 -c 1 -f -o gpurun_out/prof_knn python experiments/knn_bench.py --n 4000000 --reps 1 > gpurun_out/ncu_knn.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/knn_launches.csv python experiments/knn_bench.py --n 4000000 --reps 1 > gpurun_out/ncu_knn2.log 2>&1
cat gpurun_out/knn_tests.log gpurun_out/knn_bench.log; tail -3 gpurun_out/ncu_knn.log

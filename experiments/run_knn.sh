#!/bin/bash
# kNN on the GPU box: parity tests at every cluster size, then the micro-benchmark.
mkdir -p gpurun_out
for c in 1 2 4 8; do
  echo "== tests cluster=$c" >> gpurun_out/knn_tests.log
  DBX_KNN_CLUSTER=$c timeout 300 python -m pytest tests/test_knn_gpu.py -x -q -m gpu -k "random_768 or ragged or tensor_core" 2>&1 | tail -15 >> gpurun_out/knn_tests.log
done
echo "== full default" >> gpurun_out/knn_tests.log
timeout 600 python -m pytest tests/test_knn_gpu.py -x -q -m gpu 2>&1 | tail -15 >> gpurun_out/knn_tests.log
for c in 1 2 4 8; do
  echo "== bench cluster=$c" >> gpurun_out/knn_bench.log
  DBX_KNN_CLUSTER=$c timeout 300 python experiments/knn_bench.py --n ${KNN_N:-4000000} --reps 3 2>&1 | tail -4 >> gpurun_out/knn_bench.log
done
DBX_KNN_CLUSTER=${KNN_BEST:-4} timeout 300 python experiments/knn_bench.py --n ${KNN_N:-4000000} --reps 2 --fn l2_distance 2>&1 | tail -3 >> gpurun_out/knn_bench.log
cat gpurun_out/knn_tests.log gpurun_out/knn_bench.log

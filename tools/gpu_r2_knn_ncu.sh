#!/bin/bash
# ncu --set full of the three kNN kernels, NL=2 (default) and NL=1, at the bench's call size (128 clouds per call)
set -u
mkdir -p gpurun_out
for L in 2 1; do
  EPC_KNN_LISTS=$L timeout 900 ncu --set full --clock-control none --import-source on -k "regex:knn_bound|knn_collect|knn_slow" -s 3 -c 3 -o gpurun_out/r2_knn_L$L -f \
     python bench.py --steps 1 --warmup 1 --clouds 128 --chunk 128 --streams 1 --no-cpu-baseline --no-retrieval --no-parity > gpurun_out/ncu_r2_knn_L$L.log 2>&1
  tail -2 gpurun_out/ncu_r2_knn_L$L.log | cut -c1-200
done
ls -la gpurun_out/*.ncu-rep | tail -3

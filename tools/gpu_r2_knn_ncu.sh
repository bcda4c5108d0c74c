#!/bin/bash
# ncu --set full of the kNN kernels at the bench's call size (128 clouds per call)
set -u
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:knn_bound|knn_collect" -s 2 -c 2 -o gpurun_out/r2_knn_g8 -f \
     python bench.py --steps 1 --warmup 1 --clouds 128 --chunk 128 --streams 1 --no-cpu-baseline --no-retrieval --no-parity > gpurun_out/ncu_r2_knn_g8.log 2>&1
tail -2 gpurun_out/ncu_r2_knn_g8.log | cut -c1-200

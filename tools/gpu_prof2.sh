#!/bin/bash
# ncu --set full captures (with source) of the top kernels at the bench's real batch (128 clouds per call).
# usage: gpu_prof2.sh <tag>      -> gpurun_out/<tag>_{front,head}.ncu-rep
set -u
tag=${1:-r1b}
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --clouds 128 --chunk 128 --no-cpu-baseline --no-retrieval"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:knn_kernel|proxy_block_kernel|sort_kernel|conv_in_kernel" -s 7 -c 7 \
    -o gpurun_out/${tag}_front -f $B > gpurun_out/ncu_${tag}_front.log 2>&1
tail -1 gpurun_out/ncu_${tag}_front.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:tc_gemm|vlad_" -s 52 -c 8 \
    -o gpurun_out/${tag}_head -f $B > gpurun_out/ncu_${tag}_head.log 2>&1
tail -1 gpurun_out/ncu_${tag}_head.log | cut -c1-200
ls -la gpurun_out/${tag}_*.ncu-rep

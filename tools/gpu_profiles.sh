#!/bin/bash
# Round-end evidence: tests, smoke, bench (both arms), ncu launch list of the bench command, ncu --set full of the top kernels.
# usage: gpu_profiles.sh <tag>
set -u
tag=${1:-r1}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=line -p no:cacheprovider 2>&1 | tail -8 | cut -c1-300 | tee gpurun_out/pytest_gpu_${tag}.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee gpurun_out/smoke_${tag}.log
timeout 900 python bench.py 2> gpurun_out/bench.err | tee gpurun_out/bench_${tag}.json | cut -c1-300
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>> gpurun_out/bench.err | tee gpurun_out/bench_ref_${tag}.json | cut -c1-200
timeout 600 python bench.py --arch epc-net-l --clouds 512 --chunk 256 --no-retrieval --cpu-sample 12 2>> gpurun_out/bench.err | tee gpurun_out/bench_l_${tag}.json | cut -c1-200
echo "== ncu launch list (same command, short)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-retrieval > gpurun_out/ncu_launch.log 2>&1
tail -1 gpurun_out/ncu_launch.log | cut -c1-200
echo "== ncu --set full of the top kernels (second library call: launches 21..)"
B="python bench.py --steps 1 --warmup 1 --clouds 128 --chunk 128 --no-cpu-baseline --no-retrieval"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:knn_kernel|proxy_block_kernel|sort_kernel|conv_in_kernel" -s 7 -c 7 \
    -o gpurun_out/${tag}_front -f $B > gpurun_out/ncu_${tag}_front.log 2>&1
tail -1 gpurun_out/ncu_${tag}_front.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:tc_gemm|vlad_" -s 11 -c 11 \
    -o gpurun_out/${tag}_head -f $B > gpurun_out/ncu_${tag}_head.log 2>&1
tail -1 gpurun_out/ncu_${tag}_head.log | cut -c1-200
ls -la gpurun_out/${tag}_*.ncu-rep

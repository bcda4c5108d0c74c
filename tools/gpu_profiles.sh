#!/bin/bash
# Round-end evidence: tests, bench, ncu launch list of the bench command, ncu --set full of the top kernels.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=line -p no:cacheprovider 2>&1 | tail -8 | cut -c1-300 | tee gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 2> gpurun_out/bench.err | tee gpurun_out/bench_r1.json | cut -c1-400
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>> gpurun_out/bench.err | tee gpurun_out/bench_ref_r1.json | cut -c1-300
echo "== ncu launch list (same command, short)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-retrieval > gpurun_out/ncu_launch.log 2>&1
tail -1 gpurun_out/ncu_launch.log | cut -c1-200
echo "== ncu --set full of the top kernels"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:proxy_block_kernel|knn_kernel|tc_gemm_bres_kernel|tc_gemm_kernel" -s 8 -c 10 -o gpurun_out/top_kernels_r1 -f \
    python bench.py --steps 1 --warmup 1 --clouds 32 --chunk 32 --no-cpu-baseline --no-retrieval > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep | tail -3

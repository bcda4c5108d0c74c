#!/bin/bash
# Round-end evidence: tests, smoke, bench (both arms), ncu launch list of the bench command, ncu --set full of the top kernels.
# usage: gpu_profiles.sh <tag>
set -u
tag=${1:-r2}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=line -p no:cacheprovider 2>&1 | tail -8 | cut -c1-300 | tee gpurun_out/pytest_gpu_${tag}.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee gpurun_out/smoke_${tag}.log
timeout 900 python bench.py --steps 20 --warmup 5 2> gpurun_out/bench.err | tee gpurun_out/bench_${tag}.json | cut -c1-300
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>> gpurun_out/bench.err | tee gpurun_out/bench_ref_${tag}.json | cut -c1-200
echo "== randomized parity sweep (oracle-checked; N = 2048 / 4096 so that the fp8 head and the fp16 EPC-Net-L path are the ones exercised)"
STRESS_N=2048,4096,4096 STRESS_SEED=11 timeout 1200 python tools/gpu_stress.py 36 > gpurun_out/stress_${tag}.log 2>&1; tail -1 gpurun_out/stress_${tag}.log
echo "== ncu launch list (same command, short)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 1 --warmup 1 --clouds 512 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
tail -1 gpurun_out/ncu_launch.log | cut -c1-200
echo "== ncu --set full of the top kernels (128 clouds per call)"
B="python bench.py --steps 1 --warmup 1 --clouds 128 --batch 128 --chunk 128 --streams 1 --no-cpu-baseline --no-retrieval --no-epc-net-l --no-parity"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:knn_|proxy_block_kernel|sort_kernel|conv_in_kernel" -s 10 -c 10 \
    -o gpurun_out/${tag}_front -f $B > gpurun_out/ncu_${tag}_front.log 2>&1
tail -1 gpurun_out/ncu_${tag}_front.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:tc_gemm|vlad_|assign_vlad|sprime" -s 7 -c 7 \
    -o gpurun_out/${tag}_head -f $B > gpurun_out/ncu_${tag}_head.log 2>&1
tail -1 gpurun_out/ncu_${tag}_head.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:tc_gemm_bres" -s 0 -c 2 \
    -o gpurun_out/${tag}_colmax -f python bench.py --arch epc-net-l --steps 1 --warmup 1 --clouds 128 --batch 128 --chunk 128 --streams 1 --no-cpu-baseline --no-retrieval --no-parity > gpurun_out/ncu_${tag}_colmax.log 2>&1
tail -1 gpurun_out/ncu_${tag}_colmax.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:retr_score|select_kernel|rerank_kernel|sample_threshold|split2" -s 12 -c 6 \
    -o gpurun_out/${tag}_retrieval -f python tools/retr_prof.py > gpurun_out/ncu_${tag}_retrieval.log 2>&1
tail -1 gpurun_out/ncu_${tag}_retrieval.log | cut -c1-200
echo "== conv5 per-tile timeline (clock64 stamps of CTA 0) and the ALU-rate probe"
EPC_BRES_TIMELINE=1 timeout 200 python bench.py --steps 1 --warmup 1 --clouds 512 --no-cpu-baseline --no-retrieval --no-parity --no-epc-net-l 2>&1 | grep "conv5 tile" | head -40 > gpurun_out/${tag}_conv5_timeline.txt
tail -2 gpurun_out/${tag}_conv5_timeline.txt
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/alu_probe tools/alu_probe.cu && timeout 120 /tmp/alu_probe > gpurun_out/${tag}_alu_probe.txt; tail -2 gpurun_out/${tag}_alu_probe.txt
ls -la gpurun_out/${tag}_*.ncu-rep

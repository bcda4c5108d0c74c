#!/bin/bash
# round-2 first GPU call: microbenchmarks for the kNN redesign + the full GPU test suite (with the new batch-shape tests)
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tee gpurun_out/gpu.txt
timeout 120 tools/bin/microbench2 2>&1 | tee gpurun_out/microbench2.txt
timeout 1700 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -x 2>&1 | tail -40 | cut -c1-300 | tee gpurun_out/pytest_gpu_r2a.log

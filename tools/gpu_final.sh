#!/bin/bash
# short round-end check: GPU tests, smoke, the bench line at the driver's arguments, the reference arm
set -u
tag=${1:-r2}
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -q --tb=line -p no:cacheprovider 2>&1 | tail -4 | cut -c1-300 | tee gpurun_out/pytest_gpu_${tag}.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke_${tag}.log
timeout 200 python bench.py --steps 20 --warmup 5 2> gpurun_out/bench.err | tee gpurun_out/bench_${tag}.json | cut -c1-200
timeout 60 python bench.py --impl reference --steps 2 --warmup 1 2>> gpurun_out/bench.err | tee gpurun_out/bench_ref_${tag}.json | cut -c1-160

#!/bin/bash
# ncu --set full capture of selected kernels (regex in $1) during a tiny bench run; report -> gpurun_out/$2.ncu-rep
set -u
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$1" -s ${3:-1} -c ${4:-2} -o gpurun_out/$2 -f \
   python bench.py --steps 1 --warmup 1 --clouds 32 --no-cpu-baseline --no-retrieval > gpurun_out/ncu_$2.log 2>&1
tail -2 gpurun_out/ncu_$2.log | cut -c1-200
ls -la gpurun_out/$2.ncu-rep

#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_knn.py -m gpu -q --tb=short -p no:cacheprovider -x 2>&1 | tail -5 | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:knn_bound|knn_collect|knn_finalize" -s 3 -c 3 -o gpurun_out/r2_knn_g8b -f \
     python bench.py --steps 1 --warmup 1 --clouds 128 --chunk 128 --streams 1 --no-cpu-baseline --no-retrieval --no-parity > gpurun_out/ncu_r2_knn_g8b.log 2>&1
tail -2 gpurun_out/ncu_r2_knn_g8b.log | cut -c1-200
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-retrieval 2> gpurun_out/bench_x.err | tee gpurun_out/bench_x.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'parity',d.get('max_abs'),d.get('min_cos'))
for k,v in sorted(d['stages'].items(), key=lambda kv:-kv[1]['share'])[:3]: print('  %-16s %8.2f us/cloud  %5.1f%%'%(k,v['ms_per_cloud']*1e3,v['share']*100))
"

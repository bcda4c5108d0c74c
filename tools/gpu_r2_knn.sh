#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_knn.py -m gpu -q --tb=short -p no:cacheprovider -x 2>&1 | tail -30 | cut -c1-300 | tee gpurun_out/pytest_knn.log
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -30 | cut -c1-300 | tee gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-retrieval 2> gpurun_out/bench_x.err | tee gpurun_out/bench_x.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'parity',d.get('max_abs'),d.get('min_cos'))
for k,v in sorted(d['stages'].items(), key=lambda kv:-kv[1]['share'])[:3]: print('  %-16s %8.2f us/cloud  %5.1f%%'%(k,v['ms_per_cloud']*1e3,v['share']*100))
"
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:"knn" -c 16 --csv --log-file gpurun_out/knn_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-retrieval --no-parity > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/knn_launches.csv')) if len(r)>10 and r[0].isdigit()]
agg={}
for r in rows:
    name=r[4].split('(')[0][:44]; agg.setdefault((name,r[-3]),[]).append(float(r[-1].replace(',','')))
for k,v in agg.items(): print('%-44s %-28s n=%d mean=%.1f'%(k[0],k[1],len(v),sum(v)/len(v)))
PY

#!/bin/bash
# quick GPU iteration: parity tests + short bench (+ optional extra command in $1)
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=line -p no:cacheprovider 2>&1 | tail -25 | cut -c1-400 | tee gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench.err | tee gpurun_out/bench.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'launches',d['gpu_launches'],'clocks',d['clocks'])
for k,v in sorted(d['stages'].items(), key=lambda kv:-kv[1]['share']): print('  %-16s %8.2f us/cloud  %5.1f%%'%(k,v['us_per_cloud'],v['share']*100))
print('roofline',d['roofline'])
print('retrieval',d.get('retrieval'))
print('epc_net_l',d.get('epc_net_l'))
"
tail -3 gpurun_out/bench.err
if [ $# -ge 1 ]; then eval "$1"; fi

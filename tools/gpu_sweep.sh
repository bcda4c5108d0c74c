#!/bin/bash
# sweep an environment variable over values: gpu_sweep.sh VAR v1 v2 ...   (prints the bench stage table per value)
var=$1; shift
for v in "$@"; do
  echo "== $var=$v"
  env $var=$v timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-retrieval 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1))
for k,v in sorted(d['stages'].items(), key=lambda kv:-kv[1]['share']): print('  %-16s %8.2f us/cloud  %5.1f%%'%(k,v['ms_per_cloud']*1e3,v['share']*100))
"
done

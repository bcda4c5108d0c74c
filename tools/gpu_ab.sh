# A/B driver for one environment switch: usage: gpu_ab.sh VAR v1 v2 ... [-- extra bench args]
mkdir -p gpurun_out
VAR=$1; shift
VALS=(); while [ $# -gt 0 ] && [ "$1" != "--" ]; do VALS+=("$1"); shift; done; [ "${1:-}" = "--" ] && shift
for v in "${VALS[@]}"; do
echo "== $VAR=$v"
env $VAR=$v timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_batch_shape.py -m gpu -q --tb=line -p no:cacheprovider 2>&1 | tail -4 | cut -c1-300
env $VAR=$v timeout 600 python bench.py --steps 5 --warmup 2 --no-cpu-baseline --no-retrieval "$@" 2>gpurun_out/bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
s=d['stages']
print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'parity',d.get('max_abs'),' '.join('%s %.2f'%(k,s[k]['us_per_cloud']) for k in ('knn','proxy_block','conv5','assign_vlad','fc') if k in s))
L=d.get('epc_net_l')
if L: print('  L', round(L['value'],1), L['parity']['max_abs'], {k:round(v['us_per_cloud'],2) for k,v in L['stages'].items()})
"; tail -2 gpurun_out/bench.err
done

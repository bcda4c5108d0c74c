# stage-time A/B of one environment switch, no parity tests (for debug modes that change results on purpose)
# usage: gpu_stage_ab.sh VAR v1 v2 ... [-- extra bench args]
mkdir -p gpurun_out
VAR=$1; shift
VALS=(); while [ $# -gt 0 ] && [ "$1" != "--" ]; do VALS+=("$1"); shift; done; [ "${1:-}" = "--" ] && shift
for v in "${VALS[@]}"; do
echo "== $VAR=$v"
env $VAR=$v timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-retrieval --no-parity --no-epc-net-l "$@" 2>gpurun_out/bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
s=d['stages']
print('value',round(d['value'],1),' '.join('%s %.2f'%(k,s[k]['us_per_cloud']) for k in ('knn','proxy_block','conv5','assign_vlad','fc') if k in s), d['clocks'])
"; tail -2 gpurun_out/bench.err
done

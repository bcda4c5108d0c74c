mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout 600 python bench.py --steps 4 --warmup 2 --no-cpu-baseline --no-retrieval --no-epc-net-l --no-parity $EXTRA 2>gpurun_out/bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
s=d['stages']
print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),' '.join('%s %.2f'%(k,s[k]['us_per_cloud']) for k in ('knn','proxy_block','conv5','assign_vlad','sort') if k in s))
"; tail -2 gpurun_out/bench.err; }
EXTRA="" run EPC_HEAD_SUB=64
EXTRA="" run EPC_HEAD_SUB=128
EXTRA="--chunk 256 --batch 512" run EPC_HEAD_SUB=64
EXTRA="--chunk 256 --batch 512" run EPC_HEAD_SUB=128
EXTRA="--chunk 256 --batch 512" run EPC_HEAD_SUB=256
EXTRA="--chunk 128 --batch 384 --streams 3" run EPC_HEAD_SUB=128

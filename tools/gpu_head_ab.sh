mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout 600 python bench.py --steps 4 --warmup 2 --no-cpu-baseline --no-retrieval --no-epc-net-l --no-parity 2>gpurun_out/bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
s=d['stages']
print('value',round(d['value'],1),' '.join('%s %.2f'%(k,s[k]['us_per_cloud']) for k in ('conv5','assign_vlad','vlad_finalize') if k in s))
"; tail -2 gpurun_out/bench.err; }
run EPC_HEAD_FP8=1 EPC_HEAD_ASSIGN_CTAS=115
run EPC_HEAD_FP8=1 EPC_HEAD_ASSIGN_CTAS=125
run EPC_HEAD_FP8=1 EPC_HEAD_ASSIGN_CTAS=105 EPC_CONV5_PREFETCH=2
run EPC_HEAD_FP8=1 EPC_HEAD_ASSIGN_CTAS=105 EPC_CONV5_PREFETCH=6
run EPC_HEAD_FP8=0 EPC_CONV5_PREFETCH=4

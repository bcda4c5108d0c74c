mkdir -p gpurun_out
for f in 0 1; do
echo "== EPC_L_BF16=$f"
EPC_L_BF16=$f timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_batch_shape.py -m gpu -q --tb=line -p no:cacheprovider -k "epc-net-l or epc_net_l or golden or oracle_small or permutation or output_dim" 2>&1 | tail -4 | cut -c1-300
EPC_L_BF16=$f timeout 600 python bench.py --arch epc-net-l --steps 5 --warmup 2 --batch 512 --chunk 256 --no-cpu-baseline --no-retrieval 2>gpurun_out/bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
s=d['stages']
print('L value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'parity',d.get('max_abs'),' '.join('%s %.2f'%(k,s[k]['us_per_cloud']) for k in ('knn','proxy_block','conv5','fc') if k in s))
"; tail -2 gpurun_out/bench.err
done

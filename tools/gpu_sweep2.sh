mkdir -p gpurun_out
for sk in 2 4; do for a in 105 95; do
echo "== EPC_VLAD_SPLITK=$sk n_assign=$a"
EPC_VLAD_SPLITK=$sk EPC_HEAD_ASSIGN_CTAS=$a timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k regex:assign_vlad -s 1 -c 1 --csv --log-file gpurun_out/hf.csv python bench.py --steps 1 --warmup 1 --clouds 128 --batch 128 --chunk 128 --streams 1 --no-cpu-baseline --no-retrieval --no-epc-net-l --no-parity > /dev/null 2>&1
grep -E "dram__bytes|gpu__time" gpurun_out/hf.csv | awk -F'","' '{print $(NF-2), $(NF)}' | tr -d '"' | tr '\n' ' '; echo
EPC_VLAD_SPLITK=$sk EPC_HEAD_ASSIGN_CTAS=$a timeout 600 python bench.py --steps 4 --warmup 2 --no-cpu-baseline --no-retrieval --no-epc-net-l --no-parity 2>gpurun_out/bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
s=d['stages']
print('value',round(d['value'],1),' '.join('%s %.2f'%(k,s[k]['us_per_cloud']) for k in ('conv5','assign_vlad','vlad_finalize') if k in s))
"; tail -2 gpurun_out/bench.err
done; done
echo "== streams 3"; timeout 600 python bench.py --steps 4 --warmup 2 --streams 3 --batch 768 --no-cpu-baseline --no-retrieval --no-epc-net-l --no-parity 2>gpurun_out/bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1))"
echo "== head sub 256"; EPC_HEAD_SUB=256 timeout 600 python bench.py --steps 4 --warmup 2 --no-cpu-baseline --no-retrieval --no-epc-net-l --no-parity 2>gpurun_out/bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stages']; print('value',round(d['value'],1),' '.join('%s %.2f'%(k,s[k]['us_per_cloud']) for k in ('conv5','assign_vlad')))"

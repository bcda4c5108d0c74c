mkdir -p gpurun_out
for h in 1 2 3; do
echo "== EPC_HEAD_L2_HINTS=$h"
EPC_HEAD_L2_HINTS=$h timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:assign_vlad -c 2 --csv --log-file gpurun_out/hf_$h.csv python bench.py --steps 1 --warmup 1 --clouds 64 --batch 64 --chunk 64 --no-cpu-baseline --no-retrieval --no-epc-net-l --no-parity > /dev/null 2>&1
grep -E "dram__bytes|gpu__time|hit_rate" gpurun_out/hf_$h.csv | awk -F'","' '{print $(NF-2), $(NF)}' | tr -d '"' | head -4
done

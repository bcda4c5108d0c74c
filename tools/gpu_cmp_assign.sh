for v in 1 0; do
  if [ $v = 1 ]; then export EPC_ASSIGN_FORWARD=1; else unset EPC_ASSIGN_FORWARD; fi
  python bench.py --no-cpu-baseline --no-retrieval 2>/dev/null | tail -1 > gpurun_out/cmp_$v.json
  python - <<PY
import json
d=json.loads(open("gpurun_out/cmp_$v.json").read())
s=d["stages"]
print("forward=$v value %.0f e2e %.0f  conv5 %.3f assign %.3f vlad %.3f us/cloud" % (d["value"], d["e2e"]["value"], s["conv5"]["ms_per_cloud"]*1e3, s["assign_gemm"]["ms_per_cloud"]*1e3, s["vlad_gemm"]["ms_per_cloud"]*1e3))
PY
done

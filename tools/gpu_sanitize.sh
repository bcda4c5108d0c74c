#!/bin/bash
# compute-sanitizer passes over the hot path at small sizes: memcheck, racecheck, synccheck, initcheck.
# usage: gpu_sanitize.sh <tag>     -> gpurun_out/<tag>_sanitizer_<tool>.txt
set -u
tag=${1:-r2}
mkdir -p gpurun_out
cat > /tmp/san_small.py <<'PY'
import importlib, os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch, _data
variables = importlib.import_module("epc-net_b200.variables")
models = importlib.import_module("epc-net_b200.models")
evaluate = importlib.import_module("epc-net_b200.evaluate")
tf_util = importlib.import_module("epc-net_b200.utils.tf_util")
N = int(os.environ.get("SAN_N", 2048))
clouds = np.stack([_data.cloud(k, 50 + i, N) for i, k in enumerate(["uniform", "clustered", "zeros"])], 0)
idx, kth, cnt = tf_util.knn_graph(torch.from_numpy(clouds).cuda())
for arch in ("epc-net", "epc-net-l"):
    V = variables.synthetic_variables(arch, 5)
    params = dict(_data.default_params(arch), NUM_POINTS=N, VARIABLES=variables.VariableStore(V))
    out = models.load(arch).forward(torch.from_numpy(clouds[None]).cuda(), False, params=params)
    assert np.isfinite(out.cpu().numpy()).all()
for D, Q in ((5000, 130), (700, 40)):                      # fused candidate filter path, dense path
    db, q, _ = _data.retrieval_problem(D=D, Q=Q, seed=3)
    d, i = evaluate.retrieve_topk(db, q, 25)
    assert (i.cpu().numpy()[:, 0] >= 0).all()
torch.cuda.synchronize()
print("sanitizer workload ok")
PY
for tool in memcheck racecheck synccheck initcheck; do
    echo "== $tool"
    timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san_small.py > gpurun_out/${tag}_sanitizer_${tool}.txt 2>&1
    echo "rc=$?" >> gpurun_out/${tag}_sanitizer_${tool}.txt
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitizer workload ok|rc=" gpurun_out/${tag}_sanitizer_${tool}.txt | head -5
done

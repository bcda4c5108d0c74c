import importlib, os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch, _data
from oracle import epc_oracle
variables = importlib.import_module("epc-net_b200.variables"); models = importlib.import_module("epc-net_b200.models")
tfu = importlib.import_module("epc-net_b200.utils.tf_util")
arch = sys.argv[1] if len(sys.argv) > 1 else "epc-net"
V = variables.synthetic_variables(arch, 8)
params = dict(_data.default_params(arch), VARIABLES=variables.VariableStore(V))
normal = np.stack([_data.cloud("uniform", 300 + i, 4096) for i in range(3)], 0)
mixed = np.concatenate([normal[:1], np.zeros((1, 4096, 3), np.float32), normal[1:]], 0)
ref = epc_oracle.forward(arch, mixed[None, 1:2], V, params)
f = models.load(arch).forward
x = torch.from_numpy(mixed[None]).cuda()
# something else first, like the test-suite does (different shapes through the same workspace)
small = np.stack([_data.cloud(k, 700 + i, 512) for i, k in enumerate(["uniform", "clustered", "coarse", "duplicated", "zeros", "planar"])], 0)
tfu.knn_graph(torch.from_numpy(small).cuda())
bad = 0
for it in range(12):
    if it % 3 == 0:
        tfu.knn_graph(torch.from_numpy(small).cuda())
    b = f(x, False, params=params)[0]
    err = float(np.abs(b[1:2].cpu().numpy() - ref).max())
    bad += err > 1e-3
    print(it, "zero-cloud err %.3e" % err, flush=True)
print("bad", bad, "of 12")

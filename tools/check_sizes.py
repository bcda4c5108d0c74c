import importlib, os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch, _data
from oracle import knn_c, epc_oracle
tfu = importlib.import_module("epc-net_b200.utils.tf_util"); variables = importlib.import_module("epc-net_b200.variables"); models = importlib.import_module("epc-net_b200.models")
for N in (8192, 128, 32):
    clouds = np.stack([_data.cloud(k, 77 + i, N) for i, k in enumerate(["uniform", "coarse"])], 0)
    idx, kth, cnt = tfu.knn_graph(torch.from_numpy(clouds).cuda())
    oi, ok, oc = knn_c.knn(clouds)
    print(N, "knn exact:", np.array_equal(kth.cpu().numpy().view(np.uint32), ok.view(np.uint32)), np.array_equal(cnt.cpu().numpy(), oc), np.array_equal(idx.cpu().numpy(), oi))
    if N >= 128:
        arch = "epc-net-l"
        V = variables.synthetic_variables(arch, 9)
        params = dict(_data.default_params(arch), NUM_POINTS=N, VARIABLES=variables.VariableStore(V))
        out = models.load(arch).forward(torch.from_numpy(clouds[None]).cuda(), False, params=params).cpu().numpy()
        if N <= 2048 or True:
            ref = epc_oracle.forward(arch, clouds[None], V, params)
            print(N, arch, "max|d| %.2e" % np.abs(out - ref).max())

# run the kNN GPU tests in-process, then loop the zero-cloud scenario (diagnostic for an order-dependent failure)
import importlib, os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import pytest
rc = pytest.main(["-q", "-m", "gpu", "--tb=line", "-p", "no:cacheprovider"] + sys.argv[1:])
print("pytest rc", rc, flush=True)
import numpy as np, torch, _data
from oracle import epc_oracle
variables = importlib.import_module("epc-net_b200.variables"); models = importlib.import_module("epc-net_b200.models")
for arch in ("epc-net", "epc-net-l"):
    V = variables.synthetic_variables(arch, 8)
    params = dict(_data.default_params(arch), VARIABLES=variables.VariableStore(V))
    normal = np.stack([_data.cloud("uniform", 300 + i, 4096) for i in range(3)], 0)
    mixed = np.concatenate([normal[:1], np.zeros((1, 4096, 3), np.float32), normal[1:]], 0)
    ref = epc_oracle.forward(arch, mixed[None, 1:2], V, params)
    f = models.load(arch).forward
    x = torch.from_numpy(mixed[None]).cuda()
    errs = []
    for it in range(10):
        b = f(x, False, params=params)[0]
        errs.append(float(np.abs(b[1:2].cpu().numpy() - ref).max()))
    print(arch, " ".join("%.2e" % e for e in errs), flush=True)

"""Per-cloud descriptor error of the 18 golden clouds (tests/golden/graph_epc-net.npz) -- which kinds a head variant hurts."""
import importlib, os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch, _data
variables = importlib.import_module("epc-net_b200.variables"); models = importlib.import_module("epc-net_b200.models")
g = np.load("tests/golden/graph_epc-net.npz")
V = variables.synthetic_variables("epc-net", int(g["weight_seed"]), str(g["scope"]))
clouds = _data.golden_batch(int(g["cloud_seed"]))
params = dict(_data.default_params("epc-net"), VARIABLES=variables.VariableStore(V))
with variables.variable_scope(str(g["scope"])):
    out = models.load("epc-net").forward(torch.from_numpy(clouds[None]).cuda(), False, params=params).cpu().numpy().reshape(18, 256)
ref = g["output"].reshape(18, 256)
err = np.abs(out - ref).max(1); cos = (out * ref).sum(1) / (np.linalg.norm(out, axis=1) * np.linalg.norm(ref, axis=1))
print("err", " ".join("%.1e" % e for e in err))
print("cos", " ".join("%.6f" % c for c in cos))

#!/usr/bin/env python
"""One-off randomized parity sweep (not part of the test suite): mixed cloud kinds, sizes and batch shapes against the oracle."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, _data
from oracle import epc_oracle, knn_c
variables = importlib.import_module("epc-net_b200.variables"); models = importlib.import_module("epc-net_b200.models")
tfu = importlib.import_module("epc-net_b200.utils.tf_util")
kinds = ["uniform", "clustered", "coarse", "duplicated", "planar", "zeros", "quantised"]
rng = np.random.default_rng(int(os.environ.get("STRESS_SEED", 123)))
worst = 0.0
for trial in range(int(sys.argv[1]) if len(sys.argv) > 1 else 10):
    N = int(rng.choice([int(x) for x in os.environ.get("STRESS_N", "128,256,512,1024,2048,4096,4096").split(",")]))
    B = int(rng.integers(1, 9))
    arch = ["epc-net", "epc-net-l", "kd_epc-net"][trial % 3]
    ks = [kinds[int(rng.integers(0, len(kinds)))] for _ in range(B)]
    clouds = np.stack([_data.cloud(k, 5000 + 31 * trial + i, N) for i, k in enumerate(ks)], 0)
    idx, kth, cnt = tfu.knn_graph(torch.from_numpy(clouds).cuda())
    oi, ok, oc = knn_c.knn(clouds)
    assert np.array_equal(kth.cpu().numpy().view(np.uint32), ok.view(np.uint32)) and np.array_equal(cnt.cpu().numpy(), oc) and np.array_equal(idx.cpu().numpy(), oi), ("knn", trial, N, ks)
    V = variables.synthetic_variables(arch, 100 + trial)
    params = dict(_data.default_params(arch), NUM_POINTS=N, EMBED_CHUNK=int(rng.choice([1, 3, 32])), EMBED_STREAMS=int(rng.choice([1, 2, 3])),
                  VARIABLES=variables.VariableStore(V))
    res = models.load(arch).forward(torch.from_numpy(clouds[None]).cuda(), False, params=params)
    out = res[1] if arch.startswith("kd_") else res
    ref = epc_oracle.forward(arch, clouds[None], V, params)
    ref_out = ref[1] if isinstance(ref, tuple) else ref
    o = out.cpu().numpy().reshape(B, -1); r = np.asarray(ref_out).reshape(B, -1)
    err = float(np.abs(o - r).max()); nz = np.linalg.norm(r, axis=1) > 0
    cos = float(((o * r).sum(1) / np.maximum(np.linalg.norm(o, axis=1) * np.linalg.norm(r, axis=1), 1e-30))[nz].min()) if nz.any() else 1.0
    worst = max(worst, err)
    print("trial %d %-10s N=%4d B=%d kinds=%s  max|d|=%.2e cos=%.6f" % (trial, arch, N, B, ",".join(k[:3] for k in ks), err, cos), flush=True)
    assert err <= 1e-3 and cos >= 0.9999, (trial, arch, N, ks, err, cos)
print("stress ok, worst max|d| = %.2e" % worst)

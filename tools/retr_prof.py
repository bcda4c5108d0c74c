import importlib, sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch, _data
ev = importlib.import_module("epc-net_b200.evaluate"); lib = importlib.import_module("epc-net_b200._lib")
db, q, src = _data.retrieval_problem(D=20000, Q=3000, seed=7)
dbt, qt = torch.from_numpy(db).cuda(), torch.from_numpy(q).cuda()
for _ in range(2): ev.retrieve_topk(dbt, qt, 25)
torch.cuda.synchronize()
lib.profile_reset(); lib.profile_enable(True)
for _ in range(5): ev.retrieve_topk(dbt, qt, 25)
torch.cuda.synchronize()
print({k: round(v[0] / 5, 3) for k, v in lib.profile_read().items()})

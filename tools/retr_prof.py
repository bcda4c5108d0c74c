"""Per-stage event timing of one Oxford-scale retrieval (D = 20k, Q = 3k, k = 25) through a prepared index."""
import importlib, sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch, _data
ev = importlib.import_module("epc-net_b200.evaluate"); lib = importlib.import_module("epc-net_b200._lib")
D = int(os.environ.get("RETR_D", 20000)); Q = int(os.environ.get("RETR_Q", 3000))
db, q, src = _data.retrieval_problem(D=D, Q=Q, seed=7)
dbt, qt = torch.from_numpy(db).cuda(), torch.from_numpy(q).cuda()
index = ev.RetrievalIndex(dbt)
for _ in range(2): index.query(qt, 25)
torch.cuda.synchronize()
lib.profile_reset(); lib.profile_enable(True)
for _ in range(5): index.query(qt, 25)
torch.cuda.synchronize()
st = {k: round(v[0] / 5, 4) for k, v in lib.profile_read().items()}
lib.profile_enable(False)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): index.query(qt, 25)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print(st, "ms/call %.4f  Mq/s %.2f" % (ms, Q / ms / 1e3))

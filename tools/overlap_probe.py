# Experiment: does running two half-batches on two streams (kNN of one overlapping the HBM-bound head of the other) beat
# one stream?  Device-resident inputs, CUDA-event timing on a joining stream.
import importlib, os, sys, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch
variables = importlib.import_module("epc-net_b200.variables"); engine = importlib.import_module("epc-net_b200.engine")
import _data
arch = sys.argv[1] if len(sys.argv) > 1 else "epc-net"
store = variables.VariableStore(variables.synthetic_variables(arch, 1))
rng = np.random.default_rng(0)
def clouds(n): return torch.from_numpy(rng.uniform(-1, 1, (n, 4096, 3)).astype(np.float32)).cuda()

def run(nstreams, per, steps=12, warm=3, prio=False):
    eng = engine.Engine(arch, store, "query_triplets", dict(_data.default_params(arch), EMBED_CHUNK=per))
    xs = [[clouds(per) for _ in range(nstreams)] for _ in range(4)]
    outs = [torch.empty((per, 256), device="cuda") for _ in range(nstreams)]
    if prio:
        lo, hi = -1, 0
        streams = [torch.cuda.Stream(priority=(lo if i % 2 else hi)) for i in range(nstreams)]
    else:
        streams = [torch.cuda.Stream() for _ in range(nstreams)]
    main = torch.cuda.current_stream()
    def step(i):
        for s, x, o in zip(streams, xs[i % 4], outs):
            with torch.cuda.stream(s):
                eng.embed(x, out=o)
    for i in range(warm): step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main)
    for s in streams: s.wait_stream(main)
    for i in range(steps): step(i)
    for s in streams: main.wait_stream(s)
    e1.record(main)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("streams=%d x %3d clouds%s: %.1f clouds/s (%.3f ms/step)" % (nstreams, per, " prio" if prio else "", nstreams * per * steps / ms * 1e3, ms / steps), flush=True)

cfgs = sys.argv[2:] or ["1:128", "2:128"]
for c in cfgs:
    ns, per = c.split(":")
    run(int(ns), int(per))

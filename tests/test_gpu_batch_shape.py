"""-m gpu: parity AT THE BENCHMARKED BATCH SHAPE (BASELINE configs[1]/[2]; evaluate.py:351-452 is the 1-cloud contract
being widened).  One epc_embed call with B > 64 walks the head sub-batch loop (csrc/api.cu: `b0 += HEAD_SUB`) more than
once; bench.py times 2 x 256-cloud calls on two streams.  Every shape timed there is compared here with

  (a) the same clouds embedded in small calls (bit for bit: inference BN has no cross-sample coupling), and
  (b) the CPU oracle (oracle/epc_oracle.py, the dense-as-written restatement of models/epc-net.py:29-157) on the first and
      last cloud of every sub-batch, within the north_star tolerance (max-abs <= 1e-3 after L2, cosine >= 0.9999).
"""
import hashlib
import importlib
import os
import subprocess
import sys

import numpy as np
import pytest

import _data
from oracle import epc_oracle

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

TOL_ABS, TOL_COS = 1e-3, 0.9999
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def pkg(built_lib):
    class P:
        variables = importlib.import_module("epc-net_b200.variables")
        engine = importlib.import_module("epc-net_b200.engine")
        lib = importlib.import_module("epc-net_b200._lib")
    return P


def _clouds(n, seed, N=4096):
    rng = np.random.default_rng(seed)
    return rng.uniform(-1.0, 1.0, (n, N, 3)).astype(np.float32)      # bench.py make_clouds


def _check_rows(out, clouds, rows, arch, V, what):
    params = _data.default_params(arch)
    worst_abs, worst_cos = 0.0, 1.0
    for r in rows:
        ref = epc_oracle.forward(arch, clouds[r:r + 1][None], V, params).reshape(-1)
        got = out[r]
        worst_abs = max(worst_abs, float(np.abs(got - ref).max()))
        worst_cos = min(worst_cos, float((got * ref).sum() / (np.linalg.norm(got) * np.linalg.norm(ref))))
    assert worst_abs <= TOL_ABS and worst_cos >= TOL_COS, "%s: max-abs %.3e, min cos %.7f" % (what, worst_abs, worst_cos)
    return worst_abs, worst_cos


def _engine(pkg, arch, V, **kw):
    store = pkg.variables.VariableStore(V)
    return pkg.engine.Engine(arch, store, "query_triplets", dict(_data.default_params(arch), **kw))


def test_epc_net_single_call_130_clouds(pkg):
    """130 clouds in ONE epc_embed call: head sub-batches of 128 and 2 clouds."""
    arch = "epc-net"
    V = pkg.variables.synthetic_variables(arch, 1)
    clouds = _clouds(130, 1000)
    x = torch.from_numpy(clouds).cuda()
    one_call = _engine(pkg, arch, V, EMBED_CHUNK=130, EMBED_STREAMS=1)
    pkg.lib.launch_count_reset()
    big = one_call.embed(x)
    torch.cuda.synchronize()
    assert pkg.lib.launch_count() < 60, "130 clouds must be one library call, not several"
    small = _engine(pkg, arch, V, EMBED_CHUNK=48, EMBED_STREAMS=1).embed(x)          # calls of 48, 48, 34 clouds
    assert torch.equal(big, small), "a cloud's descriptor depends on the call it was batched into"
    a, c = _check_rows(big.cpu().numpy(), clouds, [0, 63, 64, 127, 128, 129, 31, 100], arch, V, "epc-net B=130")
    print("epc-net B=130 one call: max|d|=%.2e min cos=%.7f" % (a, c))


def test_epc_net_bench_step_two_streams(pkg):
    """bench.py's embed() call: 512 clouds as two 256-cloud library calls alternating over two streams (each walks the head
    sub-batch loop twice: 128 + 128 clouds)."""
    arch = "epc-net"
    V = pkg.variables.synthetic_variables(arch, 1)
    clouds = _clouds(512, 1097)
    x = torch.from_numpy(clouds).cuda()
    eng = _engine(pkg, arch, V, EMBED_CHUNK=256, EMBED_STREAMS=2)
    out = torch.empty((512, 256), dtype=torch.float32, device="cuda")
    for _ in range(3):                                   # repeated steps reuse the per-stream workspaces
        eng.embed(x, out=out)
    ref = _engine(pkg, arch, V, EMBED_CHUNK=32, EMBED_STREAMS=1).embed(x)
    assert torch.equal(out, ref)
    _check_rows(out.cpu().numpy(), clouds, [0, 127, 128, 255, 256, 383, 384, 511], arch, V, "epc-net 2x256 on two streams")
    ref128 = _engine(pkg, arch, V, EMBED_CHUNK=128, EMBED_STREAMS=2).embed(x)        # the round-1 bench shape
    assert torch.equal(out, ref128)


def test_epc_net_l_single_call_300_clouds(pkg):
    """EPC-Net-L (BASELINE configs[2], large batch): 300 clouds in one call."""
    arch = "epc-net-l"
    V = pkg.variables.synthetic_variables(arch, 1)
    clouds = _clouds(300, 2000)
    x = torch.from_numpy(clouds).cuda()
    big = _engine(pkg, arch, V, EMBED_CHUNK=300, EMBED_STREAMS=1).embed(x)
    small = _engine(pkg, arch, V, EMBED_CHUNK=64, EMBED_STREAMS=1).embed(x)
    assert torch.equal(big, small)
    _check_rows(big.cpu().numpy(), clouds, [0, 63, 64, 255, 256, 299, 150, 299 - 44], arch, V, "epc-net-l B=300")


_SUB_SCRIPT = r"""
import hashlib, importlib, os, sys
import numpy as np, torch
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
import _data
variables = importlib.import_module("epc-net_b200.variables")
engine = importlib.import_module("epc-net_b200.engine")
arch = "epc-net"
V = variables.synthetic_variables(arch, 6)
rng = np.random.default_rng(77)
clouds = rng.uniform(-1.0, 1.0, (10, 1024, 3)).astype(np.float32)
eng = engine.Engine(arch, variables.VariableStore(V), "query_triplets",
                    dict(_data.default_params(arch), NUM_POINTS=1024, EMBED_CHUNK=10, EMBED_STREAMS=1))
out = eng.embed(torch.from_numpy(clouds).cuda()).cpu().numpy()
print("SHA", hashlib.sha256(out.tobytes()).hexdigest())
"""


def test_head_sub_batch_loop_small_sub(pkg):
    """EPC_HEAD_SUB=4 (read once at library load, hence the subprocess): 10 clouds walk the conv5 -> assignment -> VLAD
    sub-batch loop three times (4, 4, 2).  Same bits as the default sub-batch of 128, and within tolerance of the oracle."""
    arch = "epc-net"
    V = pkg.variables.synthetic_variables(arch, 6)
    rng = np.random.default_rng(77)
    clouds = rng.uniform(-1.0, 1.0, (10, 1024, 3)).astype(np.float32)
    eng = _engine(pkg, arch, V, NUM_POINTS=1024, EMBED_CHUNK=10, EMBED_STREAMS=1)
    out = eng.embed(torch.from_numpy(clouds).cuda()).cpu().numpy()
    env = dict(os.environ, EPC_HEAD_SUB="4")
    res = subprocess.run([sys.executable, "-c", _SUB_SCRIPT % {"root": ROOT}], env=env, capture_output=True, text=True,
                         timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    sha = [l.split()[1] for l in res.stdout.splitlines() if l.startswith("SHA")][0]
    assert sha == hashlib.sha256(out.tobytes()).hexdigest(), "descriptors depend on the head sub-batch size"
    params = dict(_data.default_params(arch), NUM_POINTS=1024)
    ref = epc_oracle.forward(arch, clouds[None], V, params).reshape(10, 256)
    assert np.abs(out - ref).max() <= TOL_ABS and (out * ref).sum(-1).min() >= TOL_COS


@pytest.mark.parametrize("gain", [1000.0, 1.0e-3])
def test_fp8_head_is_range_safe(pkg, gain):
    """The head stores the per-point features as fp8 e4m3 (csrc/head_fp8.cu) with a per-cloud power-of-two scale derived
    from a bound of |H|.  conv5's BN gamma/beta times `gain` multiplies H = relu(BN(conv5)) by `gain` exactly
    (models/epc-net.py:136-139): 1000 would saturate e4m3 (max 448) and 1e-3 would flush to zero without the scale; with
    it the descriptors (invariant to the gain: per-point L2 norm, models/epc-net.py:147-148) stay within tolerance."""
    arch = "epc-net"
    V = dict(pkg.variables.synthetic_variables(arch, 12))
    for leaf in ("gamma", "beta"):
        k = "query_triplets/fastdgcnn/conv5/bn/" + leaf
        V[k] = (V[k] * gain).astype(np.float32)
    clouds = np.stack([_data.cloud(kind, 300 + i, 4096) for i, kind in enumerate(["uniform", "clustered", "duplicated", "zeros"])], 0)
    out = _engine(pkg, arch, V, EMBED_CHUNK=4, EMBED_STREAMS=1).embed(torch.from_numpy(clouds).cuda()).cpu().numpy()
    assert np.isfinite(out).all()
    _check_rows(out, clouds, [0, 1, 2, 3], arch, V, "epc-net, conv5 gain %g" % gain)

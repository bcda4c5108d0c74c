"""-m gpu: K1 (kNN graph) parity -- bit-exact against the canonical-arithmetic C oracle (oracle/knn_oracle.c),
which restates utils/tf_util.py:647-666, and against the graph-executed goldens."""
import importlib

import numpy as np
import pytest

import _data
from oracle import epc_oracle, knn_c

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

KINDS = ["uniform", "clustered", "planar", "quantised", "coarse", "duplicated", "zeros"]


@pytest.fixture(scope="module")
def tfu(built_lib):
    return importlib.import_module("epc-net_b200.utils.tf_util")


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


@pytest.mark.parametrize("arith", ["muladd", "fma"])
@pytest.mark.parametrize("N", [64, 512, 4096])
def test_knn_bit_exact_all_kinds(tfu, arith, N):
    clouds = np.stack([_data.cloud(k, 100 + i, N) for i, k in enumerate(KINDS)], 0)
    idx, kth, cnt = tfu.knn_graph(torch.from_numpy(clouds).cuda(), arith=arith)
    oi, ok, oc = knn_c.knn(clouds, arith=arith)
    assert np.array_equal(_bits(kth.cpu().numpy()), _bits(ok)), "kth (20th largest a) must match bit for bit"
    assert np.array_equal(cnt.cpu().numpy(), oc), "size of the thresholded set {j: a_ij >= kth_i}"
    assert np.array_equal(idx.cpu().numpy(), oi), "top-20 indices in tf.nn.top_k order"


@pytest.mark.parametrize("arith", ["muladd", "fma"])
def test_pruning_is_exact(tfu, arith):
    clouds = np.stack([_data.cloud(k, 300 + i, 4096) for i, k in enumerate(KINDS)], 0)
    x = torch.from_numpy(clouds).cuda()
    a = tfu.knn_graph(x, arith=arith, prune=True)
    b = tfu.knn_graph(x, arith=arith, prune=False)
    for u, v in zip(a, b):
        assert torch.equal(u, v)


def test_golden_graph_threshold_and_counts(tfu):
    """kth / count of the 18-cloud golden batch produced by executing the reference's shipped GraphDef."""
    g = np.load(_data_path("graph_epc-net.npz"))
    clouds = _data.golden_batch(int(g["cloud_seed"]))
    idx, kth, cnt = tfu.knn_graph(torch.from_numpy(clouds).cuda(), arith="muladd")
    assert np.array_equal(_bits(kth.cpu().numpy()), _bits(g["kth"]))
    assert np.array_equal(cnt.cpu().numpy(), g["count"])


def _data_path(name):
    import os
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name)


def test_dense_mask_and_distance(tfu):
    """pairwise_distance_mask / pairwise_distance return values (utils/tf_util.py:647-666, 577-596)."""
    import hashlib
    N = 512
    clouds = np.stack([_data.cloud(k, 400 + i, N) for i, k in enumerate(["uniform", "coarse", "zeros"])], 0)
    x = torch.from_numpy(clouds).cuda()
    mask = tfu.pairwise_distance_mask(x, k=20).cpu().numpy()
    ref = epc_oracle.pairwise_distance_mask(clouds)
    assert np.array_equal(mask, ref)
    dist = tfu.pairwise_distance(x).cpu().numpy()
    assert np.array_equal(_bits(dist), _bits(epc_oracle.pairwise_distance(clouds)))
    # mask hash of a full-size golden cloud
    g = np.load(_data_path("graph_epc-net-l.npz"))
    clouds = _data.golden_batch(int(g["cloud_seed"]))
    m = tfu.pairwise_distance_mask(torch.from_numpy(clouds).cuda()).cpu().numpy()
    assert hashlib.sha256(np.packbits(m.astype(np.bool_)).tobytes()).hexdigest() == str(g["mask_sha256"])


def test_tf_util_knn(tfu):
    """tf_util.knn(adj, k) == top_k(-adj) indices, ties -> lower index (utils/tf_util.py:599-610)."""
    rng = np.random.default_rng(0)
    adj = rng.standard_normal((3, 50, 333)).astype(np.float32)
    adj[0, 0, :] = 1.0                              # all ties
    adj[1, 1, 10:40] = -5.0                         # a tie block larger than k
    for k in (1, 5, 20, 32):
        got = tfu.knn(torch.from_numpy(adj).cuda(), k=k).cpu().numpy()
        assert np.array_equal(got, epc_oracle.knn(adj, k))


def test_bad_sizes_fail_loudly(tfu):
    lib_mod = importlib.import_module("epc-net_b200._lib")
    with pytest.raises(lib_mod.EpcError):
        tfu.knn_graph(torch.zeros((1, 48, 3), device="cuda"))        # N not a multiple of 32
    with pytest.raises(ValueError):
        tfu.knn_graph(torch.zeros((1, 64, 3)))                       # host tensor: no CPU path

"""CPU suite: pin the oracle (oracle/) against the goldens minted from the reference's own artefacts:
  * tests/golden/graph_*.npz  -- the shipped GraphDefs executed op by op (tests/golden/make_golden.py);
  * tests/golden/retrieval_kdtree.npz -- sklearn.neighbors.KDTree, the library evaluate.py:463,481 calls.
"""
import importlib
import os

import numpy as np
import pytest

import _data
from oracle import epc_oracle as O
from oracle import knn_c
from oracle import retrieval_oracle as R

variables = importlib.import_module("epc-net_b200.variables")
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


@pytest.mark.parametrize("arch", ["epc-net", "epc-net-l", "kd_epc-net", "kd_epc-net-l"])
def test_oracle_matches_graph_golden(arch):
    g = np.load(os.path.join(GOLDEN, "graph_%s.npz" % arch))
    scope = str(g["scope"])
    V = variables.synthetic_variables(arch, int(g["weight_seed"]), scope)
    sel = [0, 12, 15, 16, 17]           # uniform, clustered, coarse (ties), duplicated, zeros -- keeps the CPU suite short
    clouds = _data.golden_batch(int(g["cloud_seed"]))[sel]
    res, inter = O.forward(arch, clouds[None], V, _data.default_params(arch), scope=scope, return_intermediate=True)
    out = res[1] if isinstance(res, tuple) else res
    assert np.abs(out[0] - g["output"][0, sel]).max() <= 2e-6
    sp = g["sample_points"]
    scale = np.abs(g["conv5_rows"][sel]).max()
    assert np.abs(inter["conv5"][:, sp] - g["conv5_rows"][sel]).max() <= 1e-5 * scale
    assert np.abs(inter["concat"][:, sp] - g["concat_rows"][sel]).max() <= 1e-5 * np.abs(g["concat_rows"][sel]).max()
    if isinstance(res, tuple):
        assert np.abs(res[0].reshape(len(sel), 4096, 1024)[:, sp] - g["kd_feat_rows"][sel]).max() <= 1e-6


def test_c_knn_matches_graph_golden_and_numpy():
    g = np.load(os.path.join(GOLDEN, "graph_epc-net.npz"))
    clouds = _data.golden_batch(int(g["cloud_seed"]))
    idx, kth, cnt = knn_c.knn(clouds, arith="muladd")
    assert np.array_equal(_bits(kth), _bits(g["kth"]))
    assert np.array_equal(cnt, g["count"])
    assert cnt[17].min() == 4096 and cnt[:12].max() <= 21 and cnt[15].max() > 20      # zeros / uniform / coarse
    small = clouds[[0, 15, 17], :512]
    assert np.array_equal(knn_c.mask(small), O.pairwise_distance_mask(small))
    m = knn_c.mask(small)
    i2, _, c2 = knn_c.knn(small)
    assert np.array_equal(m.sum(-1).astype(np.int32), c2)
    rows = np.arange(512)
    for b in range(3):                                   # the sorted top-20 are members of the thresholded set
        assert (m[b][rows[:, None], i2[b]] == 1).all()


def test_knn_tie_order_and_fma_mode():
    z = np.zeros((1, 64, 3), np.float32)
    idx, kth, cnt = knn_c.knn(z)
    assert np.array_equal(idx[0, 5], np.arange(20)) and (cnt == 64).all()         # ties -> lower index first
    assert np.signbit(kth).all() and (kth == 0).all()                             # a = -(0) = -0.0
    c = _data.cloud("uniform", 1, 2048)[None]
    a0 = knn_c.mask(c, arith="muladd", want_a=True)[1]
    a1 = knn_c.mask(c, arith="fma", want_a=True)[1]
    assert (a0 != a1).any() and np.abs(a0 - a1).max() < 1e-5                      # the two reconstructions do differ


def test_oracle_knn_wrapper():
    adj = np.array([[[3., 1., 1., 0., 2.]]], np.float32)
    assert np.array_equal(O.knn(adj, 3), [[[3, 1, 2]]])


def test_retrieval_oracle_matches_kdtree_golden():
    g = np.load(os.path.join(GOLDEN, "retrieval_kdtree.npz"))
    db, q, src = _data.retrieval_problem(D=int(g["D"]), Q=int(g["Q"]), seed=int(g["seed"]))
    d, i = R.knn_f64(db, q, 25)
    assert np.array_equal(i, g["idx"])
    assert np.abs(d - g["dist"]).max() <= 1e-12
    try:
        kd, ki = R.kdtree_knn(db[:500], q[:20], 25)
    except ImportError:
        return
    d2, i2 = R.knn_f64(db[:500], q[:20], 25)
    assert np.array_equal(ki, i2)


def test_get_recall_restatement():
    dbv, qv, qsets = _data.retrieval_sets()
    rec, sim, opr = R.get_recall(dbv[0], qv[1], qsets[1], 0)
    assert rec.shape == (25,) and np.all(np.diff(rec) >= 0) and 0 < rec[0] <= 100 and rec[-1] <= 100
    assert 0 <= opr <= 100 and len(sim) > 0
    ave, avs, avo = R.evaluate_pairs(dbv, qv, qsets)
    assert ave.shape == (25,) and ave[0] > 50


def test_get_latent_vectors_batching_is_transparent():
    """evaluate.py:351-452: grouping into tuples and zero 'fake' clouds must not change any returned row."""
    arch, N = "epc-net-l", 128
    V = variables.synthetic_variables(arch, 2)
    p = dict(_data.default_params(arch), NUM_POINTS=N)
    data = np.stack([_data.cloud("uniform", 10 + i, N) for i in range(5)], 0)
    a = O.get_latent_vectors(arch, V, p, data, batch_num_queries=1)
    b = O.get_latent_vectors(arch, V, p, data, batch_num_queries=2, positives=1, negatives=0)
    assert a.shape == (5, 256) and np.abs(a - b).max() <= 1e-6


@pytest.mark.parametrize("arch", ["epc-net", "epc-net-l"])
def test_fp64_shadow_bounds_the_restatement(arch):
    """the fp32 restatement against its float64 shadow on the same neighbour mask -- the oracle's own rounding
    noise is two orders of magnitude below the 1e-3 descriptor tolerance the CUDA path is held to."""
    variables = importlib.import_module("epc-net_b200.variables")
    N = 256
    clouds = np.stack([_data.cloud(k, 810 + i, N) for i, k in enumerate(["uniform", "clustered", "coarse"])], 0)[None]
    V = variables.synthetic_variables(arch, 5)
    params = dict(_data.default_params(arch), NUM_POINTS=N)
    o32 = O.forward(arch, clouds, V, params)
    o64 = O.forward_f64(arch, clouds, V, params)
    assert o64.dtype == np.float64 and o32.dtype == np.float32
    assert np.abs(o32 - o64).max() <= 1e-5
    assert O.F32 is np.float32          # the shadow restores the working precision

"""-m gpu: K6 retrieval parity -- indices identical to sklearn's KDTree (golden) and to the float64 oracle."""
import importlib
import os

import numpy as np
import pytest

import _data
from oracle import retrieval_oracle as R

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def ev(built_lib):
    return importlib.import_module("epc-net_b200.evaluate")


def test_matches_kdtree_golden(ev):
    g = np.load(os.path.join(GOLDEN, "retrieval_kdtree.npz"))
    db, q, _ = _data.retrieval_problem(D=int(g["D"]), Q=int(g["Q"]), seed=int(g["seed"]))
    d, i = ev.retrieve_topk(db, q, 25)
    assert np.array_equal(i.cpu().numpy(), g["idx"]), "top-25 indices must equal KDTree(db).query(q, 25)"
    assert np.abs(d.cpu().numpy() - g["dist"]).max() <= 1e-12


@pytest.mark.parametrize("D,Q,k", [(20000, 512, 25), (1000, 33, 10), (40, 7, 25), (20, 3, 25), (5000, 100, 1)])
def test_matches_float64_oracle(ev, D, Q, k):
    db, q, _ = _data.retrieval_problem(D=D, Q=Q, seed=11)
    d, i = ev.retrieve_topk(db, q, k)
    rd, ri = R.knn_f64(db, q, k)
    kk = min(k, D)
    assert np.array_equal(i.cpu().numpy()[:, :kk], ri)
    assert np.abs(d.cpu().numpy()[:, :kk] - rd).max() <= 1e-12
    if k > D:
        assert (i.cpu().numpy()[:, D:] == -1).all()


@pytest.mark.parametrize("dim,D,Q", [(64, 6000, 200), (128, 9000, 130), (192, 5000, 77), (256, 4096, 129), (256, 4097, 128),
                                     (100, 3000, 50), (320, 5000, 40), (256, 127, 300)])
def test_dims_and_paths(ev, dim, D, Q):
    """Every scoring path: tcgen05 bf16-pair scoring with the fused candidate filter (dim % 64 == 0, D > 4096), its dense
    form (D <= 4096), and the FFMA path (other dims) -- indices equal the float64 brute force."""
    db, q, _ = _data.retrieval_problem(D=D, Q=Q, dim=dim, seed=dim + D)
    d, i = ev.retrieve_topk(db, q, 25)
    rd, ri = R.knn_f64(db, q, 25)
    assert np.array_equal(i.cpu().numpy(), ri)
    assert np.abs(d.cpu().numpy() - rd).max() <= 1e-12


@pytest.mark.parametrize("D,Q,k", [(500, 20, 40), (300, 7, 100), (50, 5, 64), (2000, 3, 33)])
def test_more_than_32_neighbours(ev, D, Q, k):
    """KDTree.query takes any k (train.py:857-869 passes num_to_take): k > 32 runs exact float64 scans, 32 per pass."""
    db, q, _ = _data.retrieval_problem(D=D, Q=Q, seed=k)
    db[7] = db[3]                                                   # an exact duplicate: ties -> lower index first, across passes
    d, i = ev.retrieve_topk(db, q, k)
    rd, ri = R.knn_f64(db, q, k)
    kk = min(k, D)
    assert np.array_equal(i.cpu().numpy()[:, :kk], ri)
    assert np.abs(d.cpu().numpy()[:, :kk] - rd).max() <= 1e-12
    if k > D:
        assert (i.cpu().numpy()[:, D:] == -1).all()


def test_index_object_equals_one_shot_calls(ev):
    """RetrievalIndex (the KDTree(database_output) object of evaluate.py:463): built once, queried repeatedly."""
    db, q, _ = _data.retrieval_problem(D=7000, Q=300, seed=2)
    index = ev.RetrievalIndex(db, id_offset=1000)
    for lo, hi, k in ((0, 300, 25), (0, 1, 1), (10, 140, 32), (299, 300, 25)):
        d, i = index.query(q[lo:hi], k)
        d0, i0 = ev.retrieve_topk(db, q[lo:hi], k, id_offset=1000)
        assert torch.equal(i, i0) and torch.equal(d, d0)
    rd, ri = R.knn_f64(db, q, 25)
    assert np.array_equal(index.query(q, 25)[1].cpu().numpy() - 1000, ri)


def test_empty_shard_is_all_padding(ev):
    """More ranks than database rows: a rank's shard has no rows (dist.shard_range) -> idx -1, dist +inf."""
    q = _data.retrieval_problem(D=64, Q=5, seed=1)[1]
    d, i = ev.retrieve_topk(np.zeros((0, 256), np.float32), q, 25, id_offset=7)
    assert (i.cpu().numpy() == -1).all() and np.isinf(d.cpu().numpy()).all()
    d, i = ev.RetrievalIndex(np.zeros((0, 256), np.float32)).query(q, 3)
    assert (i.cpu().numpy() == -1).all() and np.isinf(d.cpu().numpy()).all()


def test_more_queries_than_one_pass(ev):
    """Q above the per-pass query tile (8192): passes reuse the workspace; same rows as separate calls."""
    db, q, _ = _data.retrieval_problem(D=5000, Q=9000, seed=4)
    d, i = ev.retrieve_topk(db, q, 25)
    d1, i1 = ev.retrieve_topk(db, q[8000:], 25)
    d2, i2 = ev.retrieve_topk(db, q[:700], 25)
    assert torch.equal(i[8000:], i1) and torch.equal(d[8000:], d1)
    assert torch.equal(i[:700], i2) and torch.equal(d[:700], d2)
    rd, ri = R.knn_f64(db, q[8100:8300], 25)
    assert np.array_equal(i[8100:8300].cpu().numpy(), ri)


def test_trajectory_ordered_database(ev):
    """Oxford databases are ordered along the vehicle's route: neighbouring rows are similar, so a contiguous sample would
    be biased.  Place descriptors drift along the row index here; the result must still equal the float64 brute force."""
    rng = np.random.default_rng(8)
    D, dim = 12000, 256
    steps = rng.standard_normal((D, dim)).astype(np.float32) * 0.08
    db = np.cumsum(steps, 0) + rng.standard_normal((1, dim)).astype(np.float32)
    db = (db / np.linalg.norm(db, axis=1, keepdims=True)).astype(np.float32)
    src = rng.permutation(D)[:400]
    q = db[src] + 0.02 * rng.standard_normal((400, dim)).astype(np.float32)
    q = (q / np.linalg.norm(q, axis=1, keepdims=True)).astype(np.float32)
    d, i = ev.retrieve_topk(db, q, 25)
    rd, ri = R.knn_f64(db, q, 25)
    assert np.array_equal(i.cpu().numpy(), ri)


def test_candidate_region_overflow_falls_back_exactly(ev):
    """A database whose sampled rows (every stride-th 128-row tile) are all far from the queries while every other row is
    near: the sample threshold admits nearly every row, the candidate regions overflow, and the query must take the exact
    float64 scan -- never a truncated candidate list."""
    rng = np.random.default_rng(5)
    D, dim = 6000, 64
    centre = rng.standard_normal((1, dim)).astype(np.float32)
    db = centre + 0.01 * rng.standard_normal((D, dim)).astype(np.float32)
    n_tiles = (D + 255) // 256 * 2                                # retrieval.cu: 16 sample tiles of 128 rows, every stride-th
    stride = n_tiles // 16
    for t in range(16):
        lo = t * stride * 128
        db[lo:lo + 128] = 5.0 * rng.standard_normal((len(db[lo:lo + 128]), dim)).astype(np.float32)
    q = centre + 0.01 * rng.standard_normal((6, dim)).astype(np.float32)
    d, i = ev.retrieve_topk(db, q, 25)
    rd, ri = R.knn_f64(db, q, 25)
    assert np.array_equal(i.cpu().numpy(), ri)


def test_near_ties_take_the_exact_fallback(ev):
    """Database rows that differ below fp32 scoring resolution: the float64 re-rank / exact fallback must still
    return the float64 order, ties -> lower index."""
    rng = np.random.default_rng(0)
    base = rng.standard_normal((1, 256)).astype(np.float32)
    base /= np.linalg.norm(base)
    db = np.repeat(base, 300, 0)
    db[:, 0] += (np.arange(300, dtype=np.float32) % 7) * 1e-7       # many near-identical rows, some exactly equal
    db = np.concatenate([db, rng.standard_normal((500, 256)).astype(np.float32) / 16], 0)
    q = base + 1e-3 * rng.standard_normal((9, 256)).astype(np.float32)
    d, i = ev.retrieve_topk(db, q, 25)
    rd, ri = R.knn_f64(db, q, 25)
    assert np.array_equal(i.cpu().numpy(), ri)


def test_sharded_merge_is_shard_count_invariant(ev):
    dist_mod = importlib.import_module("epc-net_b200.dist")
    db, q, _ = _data.retrieval_problem(D=3001, Q=64, seed=5)
    d0, i0 = ev.retrieve_topk(db, q, 25)
    for world in (2, 3, 8):
        ds, is_ = [], []
        for r in range(world):
            s, e = dist_mod.shard_range(len(db), r, world)
            d, i = ev.retrieve_topk(db[s:e], q, 25, id_offset=s)
            ds.append(d)
            is_.append(i)
        md, mi = dist_mod.cuda_merge(torch.stack(ds), torch.stack(is_))
        assert torch.equal(mi, i0) and torch.equal(md, d0), world
        # the packed layout one all-gather delivers: [world, 2, Q, k], ids as float64 bit patterns
        g = torch.stack([torch.stack([d, i.view(torch.float64)]) for d, i in zip(ds, is_)])
        md, mi = dist_mod.cuda_merge(g[:, 0], g[:, 1].view(torch.int64))
        assert torch.equal(mi, i0) and torch.equal(md, d0), world
        md, mi = dist_mod._merge_strided(g, world, 64, 25)
        assert torch.equal(mi, i0) and torch.equal(md, d0), world


def test_get_recall_flow(ev):
    """evaluate.get_recall / evaluate()'s pair loop against the oracle restatement of evaluate.py:305-334, 455-537."""
    dbv, qv, qsets = _data.retrieval_sets()
    ev.DATABASE_VECTORS, ev.QUERY_VECTORS, ev.QUERY_SETS = dbv, qv, qsets
    tot = np.zeros(25)
    sims, oprs, cnt = [], [], 0
    for m in range(len(qsets)):
        for n in range(len(qsets)):
            if m == n:
                continue
            rec, sim, opr, for_plot = ev.get_recall(None, None, m, n)
            orec, osim, oopr = R.get_recall(dbv[m], qv[n], qsets[n], m)
            assert np.allclose(rec, orec) and opr == oopr and np.allclose(sim, osim)
            tot += rec
            cnt += 1
            sims += list(sim)
            oprs.append(opr)
    ave, avs, avo = R.evaluate_pairs(dbv, qv, qsets)
    assert np.allclose(tot / cnt, ave) and np.isclose(np.mean(sims), avs) and np.isclose(np.mean(oprs), avo)


def test_get_random_hard_negatives_matches_kdtree(ev):
    """train.py:857-869 through the GPU k-NN: same picks as the reference's per-query KDTree."""
    from sklearn.neighbors import KDTree
    rng = np.random.default_rng(11)
    latent = rng.standard_normal((4000, 256)).astype(np.float32)
    latent /= np.linalg.norm(latent, axis=1, keepdims=True)
    random_negs = rng.choice(4000, 2000, replace=False)
    q = latent[17] + 0.1 * rng.standard_normal(256).astype(np.float32)
    got = ev.get_random_hard_negatives(q, random_negs.tolist(), 10, latent)
    _, ind = KDTree(latent[random_negs]).query(np.array([q]), k=10)
    assert got == np.squeeze(random_negs[ind[0]]).tolist()


def test_query_radius_matches_kdtree(ev):
    """generating_queries/generate_test_sets.py:95-104: KDTree(...).query_radius(coor, r=25) on UTM-sized float64 coordinates."""
    from sklearn.neighbors import KDTree
    rng = np.random.default_rng(3)
    db = np.stack([5735000.0 + rng.uniform(0, 600, 500), 620000.0 + rng.uniform(0, 600, 500)], 1)
    q = np.stack([5735000.0 + rng.uniform(0, 600, 130), 620000.0 + rng.uniform(0, 600, 130)], 1)
    q[0] = db[7]                                                     # an exact hit
    got = ev.query_radius(db, q, 25)
    ref = KDTree(db).query_radius(q, r=25)
    assert len(got) == len(ref)
    for g, r in zip(got, ref):
        assert np.array_equal(g, np.sort(r))
    assert ev.query_radius(db, q[:0], 25) == []

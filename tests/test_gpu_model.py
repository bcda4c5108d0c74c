"""-m gpu: descriptor parity of MODEL.forward through the plugin surface (-> epc_embed in libepc_b200).

Tolerance (BASELINE.json north_star): max-abs <= 1e-3 after L2-normalisation and cosine >= 0.9999."""
import importlib
import os

import numpy as np
import pytest

import _data
from oracle import epc_oracle, knn_c

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

TOL_ABS, TOL_COS = 1e-3, 0.9999
ARCHS = ["epc-net", "epc-net-l", "kd_epc-net", "kd_epc-net-l"]
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def pkg(built_lib):
    class P:
        variables = importlib.import_module("epc-net_b200.variables")
        models = importlib.import_module("epc-net_b200.models")
        loupe = importlib.import_module("epc-net_b200.loupe")
        tf_util = importlib.import_module("epc-net_b200.utils.tf_util")
        evaluate = importlib.import_module("epc-net_b200.evaluate")
        engine = importlib.import_module("epc-net_b200.engine")
    return P


def _check_desc(out, ref, what=""):
    out = np.asarray(out, np.float32).reshape(-1, ref.shape[-1])
    ref = np.asarray(ref, np.float32).reshape(-1, ref.shape[-1])
    err = np.abs(out - ref).max()
    cos = ((out * ref).sum(-1) / np.maximum(np.linalg.norm(out, axis=-1) * np.linalg.norm(ref, axis=-1), 1e-30))
    # the all-zero fake cloud yields a legitimate all-zero EPC-Net-L descriptor only if every FC output is <= 0; skip cos there
    nz = np.linalg.norm(ref, axis=-1) > 0
    assert err <= TOL_ABS, "%s max-abs %.3e > %.0e" % (what, err, TOL_ABS)
    assert cos[nz].min() >= TOL_COS, "%s min cosine %.6f" % (what, cos[nz].min())
    return float(err), float(cos[nz].min())


@pytest.mark.parametrize("arch", ARCHS)
def test_forward_matches_graph_golden(pkg, arch):
    """18 clouds x 4096 points (incl. tie-heavy, duplicated and all-zero clouds) against the descriptors obtained
    by executing the reference's shipped GraphDef (tests/golden/make_golden.py)."""
    g = np.load(os.path.join(GOLDEN, "graph_%s.npz" % arch))
    scope = str(g["scope"])
    V = pkg.variables.synthetic_variables(arch, int(g["weight_seed"]), scope)
    clouds = _data.golden_batch(int(g["cloud_seed"]))
    params = dict(_data.default_params(arch), VARIABLES=pkg.variables.VariableStore(V))
    outer, inner = (scope.split("/", 1) + [None])[:2] if "/" in scope else (None, scope)
    model = pkg.models.load(arch)
    x = torch.from_numpy(clouds[None]).cuda()
    if outer:
        with pkg.variables.variable_scope(outer), pkg.variables.variable_scope(inner):
            res = model.forward(x, False, params=params)
    else:
        with pkg.variables.variable_scope(scope):
            res = model.forward(x, False, params=params)
    if arch.startswith("kd_"):
        feat, out = res
        feat = feat.cpu().numpy().reshape(18, 4096, 1024)[:, g["sample_points"], :]
        # unit-norm per-point descriptors: same tolerance as the global descriptor
        assert np.abs(feat - g["kd_feat_rows"]).max() <= TOL_ABS, "KD per-point features (models/kd_epc-net.py:158)"
        num = (feat * g["kd_feat_rows"]).sum(-1)
        den = np.linalg.norm(feat, axis=-1) * np.linalg.norm(g["kd_feat_rows"], axis=-1)
        assert (num[den > 0] / den[den > 0]).min() >= TOL_COS
    else:
        out = res
    assert tuple(out.shape) == (1, 18, 256)
    err, cos = _check_desc(out.cpu().numpy(), g["output"], arch)
    print("%s vs graph golden: max|d|=%.2e min cos=%.7f" % (arch, err, cos))


@pytest.mark.parametrize("arch", ["epc-net", "epc-net-l"])
@pytest.mark.parametrize("arith", ["muladd", "fma"])
def test_forward_matches_oracle_small(pkg, arch, arith):
    N = 512
    kinds = ["uniform", "clustered", "coarse", "duplicated", "zeros", "planar"]
    clouds = np.stack([_data.cloud(k, 700 + i, N) for i, k in enumerate(kinds)], 0).reshape(2, 3, N, 3)
    V = pkg.variables.synthetic_variables(arch, 21)
    params = dict(_data.default_params(arch), NUM_POINTS=N, KNN_ARITH=arith, VARIABLES=pkg.variables.VariableStore(V))
    out = pkg.models.load(arch).forward(torch.from_numpy(clouds).cuda(), False, params=params)
    mask = knn_c.mask(clouds.reshape(6, N, 3), arith=arith)
    ref = epc_oracle.forward(arch, clouds, V, params, mask=mask)
    assert tuple(out.shape) == (2, 3, 256)
    _check_desc(out.cpu().numpy(), ref, "%s/%s" % (arch, arith))
    # the float64 shadow of the same graph on the same mask: the fp32 restatement's own noise (~1e-6) is not what the
    # tolerance is spent on
    ref64 = epc_oracle.forward_f64(arch, clouds, V, params, mask=mask)
    assert np.abs(ref - ref64).max() <= 1e-5
    _check_desc(out.cpu().numpy(), ref64.astype(np.float32), "%s/%s vs fp64 shadow" % (arch, arith))


@pytest.mark.parametrize("arch,dim", [("epc-net", 128), ("epc-net", 512), ("epc-net", 64), ("epc-net-l", 128), ("epc-net-l", 512)])
def test_feature_output_dim_other_than_256(pkg, arch, dim):
    """FEATURE_OUTPUT_DIM (configs/*.yaml: 256) is a free parameter of the reference (loupe.py:233-247 output_dim,
    models/epc-net-l.py:95): multiples of 64 run on the tensor-core FC paths."""
    N = 512
    clouds = np.stack([_data.cloud(k, 820 + i, N) for i, k in enumerate(["uniform", "clustered", "coarse"])], 0)[None]
    V = pkg.variables.synthetic_variables(arch, 33, output_dim=dim)
    params = dict(_data.default_params(arch), NUM_POINTS=N, FEATURE_OUTPUT_DIM=dim, VARIABLES=pkg.variables.VariableStore(V))
    out = pkg.models.load(arch).forward(torch.from_numpy(clouds).cuda(), False, params=params)
    ref = epc_oracle.forward(arch, clouds, V, params)
    assert tuple(out.shape) == (1, 3, dim)
    _check_desc(out.cpu().numpy(), ref, "%s output_dim=%d" % (arch, dim))


def test_batch_independence_and_determinism(pkg):
    """Inference BN has no cross-sample coupling: a cloud's descriptor is bit-identical alone, in a batch, at any
    chunking, and run to run (this is what lets get_latent_vectors batch where the reference runs 1 cloud/sess.run)."""
    arch = "epc-net"
    V = pkg.variables.synthetic_variables(arch, 3)
    clouds = np.stack([_data.cloud("uniform", 900 + i, 4096) for i in range(5)], 0)
    x = torch.from_numpy(clouds).cuda()
    store = pkg.variables.VariableStore(V)
    e_big = pkg.engine.Engine(arch, store, "query_triplets", dict(_data.default_params(arch), EMBED_CHUNK=8))
    e_one = pkg.engine.Engine(arch, store, "query_triplets", dict(_data.default_params(arch), EMBED_CHUNK=1))
    a = e_big.embed(x)
    b = e_one.embed(x)
    c = e_big.embed(x[2:3])
    assert torch.equal(a, b) and torch.equal(a[2:3], c) and torch.equal(a, e_big.embed(x))


def test_fp16_range_fallback_is_per_cloud(pkg):
    """The ProxyConv chain runs in fp16 and re-runs in fp32 every cloud whose activations left the fp16 range (the all-zero
    "fake" clouds of evaluate.py:425-430 grow by N/20 per block).  The fallback must (a) reproduce the oracle for the
    degenerate cloud and (b) leave the other clouds of the batch bit-identical to a batch without it."""
    for arch in ("epc-net", "epc-net-l"):
        V = pkg.variables.synthetic_variables(arch, 8)
        params = dict(_data.default_params(arch), VARIABLES=pkg.variables.VariableStore(V))
        normal = np.stack([_data.cloud("uniform", 300 + i, 4096) for i in range(3)], 0)
        mixed = np.concatenate([normal[:1], np.zeros((1, 4096, 3), np.float32), normal[1:]], 0)
        f = pkg.models.load(arch).forward
        a = f(torch.from_numpy(normal[None]).cuda(), False, params=params)[0]
        b = f(torch.from_numpy(mixed[None]).cuda(), False, params=params)[0]
        assert torch.equal(a, b[[0, 2, 3]]), arch
        ref = epc_oracle.forward(arch, mixed[None, 1:2], V, params)
        _check_desc(b[1:2].cpu().numpy(), ref, arch + " zero cloud")


def test_engine_cache_is_keyed_by_store_identity_not_address(pkg):
    """A collected VariableStore's address is reused by the next one: the plugin must not serve the engine (weights) of
    the dead store.  Each round builds a fresh store with different weights at (very likely) the same address."""
    import gc
    arch = "epc-net-l"
    x = torch.from_numpy(np.stack([_data.cloud("uniform", 40, 512)], 0)[None]).cuda()
    outs = []
    for seed in (1, 2, 3, 1):
        store = pkg.variables.VariableStore(pkg.variables.synthetic_variables(arch, seed))
        params = dict(_data.default_params(arch), NUM_POINTS=512, VARIABLES=store)
        got = pkg.models.load(arch).forward(x, False, params=params)[0]
        want = pkg.engine.Engine(arch, store, "query_triplets", params).embed(x[0])
        assert torch.equal(got, want), "seed %d served by a stale engine" % seed
        outs.append(got.cpu().numpy())
        del store, params
        gc.collect()
    assert not np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[3])


def test_point_permutation_invariance(pkg):
    """Full-size property: descriptors do not depend on the order of the points of a cloud (kNN sets, max-pool and
    VLAD sums are permutation invariant; only fp32 summation order moves)."""
    rng = np.random.default_rng(5)
    for arch in ("epc-net", "epc-net-l"):
        V = pkg.variables.synthetic_variables(arch, 4)
        params = dict(_data.default_params(arch), VARIABLES=pkg.variables.VariableStore(V))
        c = np.stack([_data.cloud("uniform", 40, 4096), _data.cloud("clustered", 41, 4096)], 0)
        p = np.stack([c[0][rng.permutation(4096)], c[1][rng.permutation(4096)]], 0)
        f = pkg.models.load(arch).forward
        a = f(torch.from_numpy(c[None]).cuda(), False, params=params).cpu().numpy()
        b = f(torch.from_numpy(p[None]).cuda(), False, params=params).cpu().numpy()
        assert np.abs(a - b).max() <= 2e-5, arch


def test_loupe_interface(pkg):
    """loupe.G_VLAD / NetVLAD forward on caller-provided features (loupe.py:233-333, 119-214)."""
    rng = np.random.default_rng(1)
    B, N = 3, 256
    X = rng.standard_normal((B * N, 1024)).astype(np.float32)
    X = np.maximum(X, 0)
    X /= np.linalg.norm(X, axis=1, keepdims=True)
    for pooling, groups in (("G_VLAD", 4), ("NetVLAD", 1)):
        V = pkg.variables.synthetic_variables("epc-net", 8, "query_triplets", pooling=pooling)
        store = pkg.variables.VariableStore(V)
        with pkg.variables.variable_scope("query_triplets"), pkg.variables.variable_scope("VLAD"):
            if pooling == "G_VLAD":
                layer = pkg.loupe.G_VLAD(feature_size=1024, max_samples=N, cluster_size=64, output_dim=256, groups=4,
                                         gating=True, add_batch_norm=True, is_training=False)
            else:
                layer = pkg.loupe.NetVLAD(feature_size=1024, max_samples=N, cluster_size=64, output_dim=256,
                                          gating=True, add_batch_norm=True, is_training=False)
            layer.variables = store
            out = layer.forward(torch.from_numpy(X).cuda()).cpu().numpy()
        ref = epc_oracle.vlad_forward(X, V, "query_triplets/VLAD/", N, 64, 256, groups, True, pooling)
        assert out.shape == (B, 256)
        scale = np.abs(ref).max()
        assert np.abs(out - ref).max() <= 1e-3 * scale, pooling
    with pytest.raises(NotImplementedError):
        pkg.loupe.G_VLAD(1024, N, 64, 256, is_training=True).forward(torch.from_numpy(X).cuda())


def test_tf_util_layers(pkg):
    """conv1d / fully_connected / max_pool2d wrappers (utils/tf_util.py:52-107, 310-346, 349-372)."""
    rng = np.random.default_rng(2)
    V = pkg.variables.synthetic_variables("epc-net-l", 9)
    store = pkg.variables.VariableStore(V)
    x = rng.standard_normal((2, 300, 64)).astype(np.float32)
    with pkg.variables.variable_scope("query_triplets"), pkg.variables.variable_scope("fastdgcnn"):
        y = pkg.tf_util.conv1d(torch.from_numpy(x).cuda(), 64, 1, padding="VALID", stride=1, bn=True, is_training=False,
                               scope="conv1_a", variables_store=store).cpu().numpy()
    ref = epc_oracle.conv1d(x, V, "query_triplets/fastdgcnn/conv1_a")
    assert np.abs(y - ref).max() <= 1e-4 * max(1.0, np.abs(ref).max())
    g = rng.standard_normal((5, 1024)).astype(np.float32)
    with pkg.variables.variable_scope("query_triplets"), pkg.variables.variable_scope("VLAD"):
        o = pkg.tf_util.fully_connected(torch.from_numpy(g).cuda(), 256, bn=True, is_training=False, scope="fc1",
                                        variables_store=store).cpu().numpy()
    ref = epc_oracle.fully_connected(g, V, "query_triplets/VLAD/fc1")
    assert np.abs(o - ref).max() <= 1e-4 * max(1.0, np.abs(ref).max())
    h = rng.standard_normal((3, 128, 1, 1024)).astype(np.float32)
    mp = pkg.tf_util.max_pool2d(torch.from_numpy(h).cuda(), [128, 1], padding="VALID", scope="maxpool").cpu().numpy()
    assert np.array_equal(mp, h.max(axis=1, keepdims=True))


def test_plugin_surface_errors(pkg):
    m = pkg.models.load("epc-net")
    ph = m.placeholder_inputs(1, 1, 4096, 3)
    assert tuple(ph.shape) == (1, 1, 4096, 3) and ph.is_cuda
    V = pkg.variables.synthetic_variables("epc-net", 1)
    params = dict(_data.default_params("epc-net"), VARIABLES=pkg.variables.VariableStore(V))
    with pytest.raises(NotImplementedError):
        m.forward(ph, True, params=params)
    with pytest.raises(KeyError):
        m.forward(ph, False, params={"KNN": 20})
    empty = m.forward(torch.zeros((1, 0, 4096, 3), device="cuda"), False, params=params)   # evaluate.py feeds P=0 tensors
    assert tuple(empty.shape) == (1, 0, 256)


def test_get_latent_vectors_host_path(pkg):
    """evaluate.get_latent_vectors on host arrays == oracle restatement of evaluate.py:351-452 (ragged tail included)."""
    arch = "epc-net-l"
    N = 256
    V = pkg.variables.synthetic_variables(arch, 30)
    params = dict(_data.default_params(arch), NUM_POINTS=N, EMBED_CHUNK=4, VARIABLES=pkg.variables.VariableStore(V))
    data = np.stack([_data.cloud("uniform", 1200 + i, N) for i in range(11)], 0)
    ops = {"MODEL": pkg.models.load(arch), "params": params}
    got = pkg.evaluate.get_latent_vectors(None, ops, {i: {} for i in range(11)}, data)
    ref = epc_oracle.get_latent_vectors(arch, V, params, data, batch_num_queries=3, positives=0, negatives=0)
    assert got.shape == (11, 256)
    _check_desc(got, ref, "get_latent_vectors")
    assert pkg.evaluate.get_latent_vectors(None, ops, {}, np.zeros((0, N, 3), np.float32)).shape == (0, 256)
    # chunks alternate over EMBED_STREAMS side streams (default 2): same bits on one stream and on three, called twice
    for streams in (1, 3):
        ops_s = {"MODEL": ops["MODEL"], "params": dict(params, EMBED_STREAMS=streams)}
        for _ in range(2):
            assert np.array_equal(pkg.evaluate.get_latent_vectors(None, ops_s, {i: {} for i in range(11)}, data), got), streams


def test_train_side_callers(pkg):
    """train.py:857-965 mirrors: same rows as the evaluate path, the reference's result shapes (flat vector for one cloud,
    empty array for none) and hard negatives picked from the cached descriptors."""
    from sklearn.neighbors import KDTree
    train = importlib.import_module("epc-net_b200.train")
    arch, N = "epc-net-l", 256
    V = pkg.variables.synthetic_variables(arch, 12)
    params = dict(_data.default_params(arch), NUM_POINTS=N, VARIABLES=pkg.variables.VariableStore(V))
    ops = {"MODEL": pkg.models.load(arch), "params": params}
    data = np.stack([_data.cloud("uniform", 2100 + i, N) for i in range(9)], 0)
    train.train_data = data
    allv = train.get_latent_vectors(None, ops, {i: {} for i in range(9)})
    assert allv.shape == (9, 256)
    assert np.array_equal(allv, pkg.evaluate.get_latent_vectors(None, ops, {i: {} for i in range(9)}, data))
    one = train.get_latent_vectors(None, ops, {0: {}})
    assert one.shape == (256,) and np.array_equal(one, allv[0])
    assert train.get_latent_vectors(None, ops, {}).shape == (0,)
    train.TRAINING_LATENT_VECTORS = allv
    negs = [8, 1, 5, 3, 7, 2]
    got = train.get_random_hard_negatives(allv[0], negs, 3)
    _, ind = KDTree(allv[negs]).query(np.array([allv[0]]), k=3)
    assert got == np.squeeze(np.array(negs)[ind[0]]).tolist()


@pytest.mark.parametrize("arch,scope", [("epc-net", "query_triplets"), ("kd_epc-net-l", "student/query_triplets")])
def test_checkpoint_restore_then_forward(pkg, arch, scope, tmp_path):
    """evaluate.py:262-266 (saver.restore) -> forward: weights written as a TF V2 bundle, restored by name through
    VariableStore.restore (tf_bundle.py) and embedded, give the very descriptors of the in-memory store."""
    tf_bundle = importlib.import_module("epc-net_b200.tf_bundle")
    V = pkg.variables.synthetic_variables(arch, 17, scope)
    prefix = str(tmp_path / "model_epoch1_iter101.ckpt")
    extra = dict(V)
    extra["Variable"] = np.array(101, np.int32)                         # global step + an Adam slot, as in the shipped bundles
    extra[next(iter(V)) + "/Adam"] = np.zeros_like(V[next(iter(V))])
    tf_bundle.write_checkpoint(prefix, extra)
    restored = pkg.variables.VariableStore()
    restored.restore(prefix)
    direct = pkg.variables.VariableStore(V)
    clouds = np.stack([_data.cloud(k, 60 + i, 1024) for i, k in enumerate(["uniform", "clustered", "coarse"])], 0)
    x = torch.from_numpy(clouds[None]).cuda()
    outer, inner = scope.split("/", 1) if "/" in scope else (None, scope)
    outs = []
    for store in (restored, direct):
        params = dict(_data.default_params(arch), NUM_POINTS=1024, VARIABLES=store)
        if outer:
            with pkg.variables.variable_scope(outer), pkg.variables.variable_scope(inner):
                res = pkg.models.load(arch).forward(x, False, params=params)
        else:
            with pkg.variables.variable_scope(scope):
                res = pkg.models.load(arch).forward(x, False, params=params)
        outs.append(res[1] if arch.startswith("kd_") else res)
    assert torch.equal(outs[0], outs[1])
    ref = epc_oracle.forward(arch, clouds[None], V, dict(_data.default_params(arch), NUM_POINTS=1024), scope=scope)
    ref = ref[1] if isinstance(ref, tuple) else ref
    _check_desc(outs[0].cpu().numpy(), ref, arch + " restored checkpoint")

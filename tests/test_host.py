"""CPU suite: host logic -- variable store / checkpoint names, TF bundle reader, plugin surface, sharding (gloo x2)."""
import importlib
import os
import socket
import sys
import tempfile

import numpy as np
import pytest

import _data

variables = importlib.import_module("epc-net_b200.variables")
tf_bundle = importlib.import_module("epc-net_b200.tf_bundle")
dist_mod = importlib.import_module("epc-net_b200.dist")
REF = "/root/reference"


def test_variable_specs_sizes():
    # parameter counts of SURVEY.md section 6 (trainable = weights, biases, beta, gamma, VLAD matrices)
    def trainable(arch):
        n = 0
        for k, s in variables.variable_specs(arch).items():
            if "ExponentialMovingAverage" in k or "moving_" in k:
                continue
            n += int(np.prod(s))
        return n
    assert trainable("epc-net") == 4704832
    assert trainable("epc-net-l") == 418880
    assert trainable("kd_epc-net-l") == 418880
    s = variables.variable_specs("kd_epc-net-l", "student/query_triplets")
    assert "student/query_triplets/BACKBONE/conv5/weights" in s and s["student/query_triplets/BACKBONE/conv5/weights"] == (1, 128, 1024)
    assert variables.variable_specs("epc-net", pooling="NetVLAD")["query_triplets/VLAD/hidden1_weights"] == (65536, 256)


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted (GPU box)")
@pytest.mark.parametrize("arch,index,scope", [
    ("epc-net", "exp/epc-net/saved_model/model_epoch22_iter18101.ckpt.index", "query_triplets"),
    ("epc-net-l", "exp/epc-net-l/saved_model/model_epoch13_iter18101.ckpt.index", "query_triplets"),
    ("kd_epc-net-l", "exp/epc-net-l-d/saved_model/student_model_epoch20_iter18101.ckpt.index", "student/query_triplets"),
    ("kd_epc-net", "exp/epc-net-l-d/transfer_teacher/teacher_model_epoch22_iter18101.ckpt.index", "teacher/query_triplets"),
])
def test_names_and_shapes_match_shipped_checkpoints(arch, index, scope):
    entries = tf_bundle.read_index(os.path.join(REF, index))
    have = {k: tuple(e["shape"]) for k, e in entries.items()
            if "Adam" not in k and not k.endswith("Variable") and "beta1_power" not in k and "beta2_power" not in k}
    spec = variables.variable_specs(arch, scope)
    assert set(have) == set(spec)
    assert all(tuple(spec[k]) == have[k] for k in spec)


def test_tf_bundle_round_trip_and_restore():
    V = variables.synthetic_variables("epc-net-l", 3)
    with tempfile.TemporaryDirectory() as d:
        prefix = os.path.join(d, "model_epoch1_iter1.ckpt")
        extra = dict(V)
        extra["Variable"] = np.array(7, np.int32)
        extra["query_triplets/fastdgcnn/conv1/weights/Adam"] = np.zeros((1, 3, 64), np.float32)
        tf_bundle.write_checkpoint(prefix, extra)
        got = tf_bundle.read_checkpoint(prefix)
        assert "query_triplets/fastdgcnn/conv1/weights/Adam" not in got and int(got["Variable"]) == 7
        store = variables.VariableStore()
        store.restore(prefix)
        assert set(store.keys()) == set(V.keys())
        assert all(np.array_equal(store[k], V[k]) for k in V)
        with pytest.raises(ValueError):
            tf_bundle.read_index(prefix + ".data-00000-of-00001")
    with pytest.raises(KeyError):
        variables.VariableStore()["query_triplets/nope"]


def test_variable_scope_nesting():
    assert variables.current_scope("dflt") == "dflt"
    with variables.variable_scope("student"), variables.variable_scope("query_triplets") as s:
        assert s == "student/query_triplets" == variables.current_scope()
    assert variables.current_scope() == ""


def test_plugin_modules_expose_reference_surface():
    models = importlib.import_module("epc-net_b200.models")
    import inspect
    for arch in models.ARCHS:
        m = models.load(arch)
        assert list(inspect.signature(m.forward).parameters) == ["point_cloud", "is_training", "bn_decay", "params"]
        assert list(inspect.signature(m.placeholder_inputs).parameters) == ["batch_num_queries", "num_pointclouds_per_query",
                                                                             "num_point", "input_dim"]
    with pytest.raises(ImportError):
        models.load("pointnetvlad")
    loupe = importlib.import_module("epc-net_b200.loupe")
    g = loupe.G_VLAD(feature_size=1024, max_samples=4096, cluster_size=64, output_dim=256, groups=4, gating=True,
                     add_batch_norm=True, is_training=False)
    assert (g.groups, g.cluster_size, g.max_samples) == (4, 64, 4096)
    with pytest.raises(NotImplementedError):
        loupe.PoolingBaseModel(1024, 4096, 64, 256).forward(None)
    # the alias module gives the very same objects
    import epc_net_b200
    from epc_net_b200 import variables as v2
    assert v2 is variables and epc_net_b200.__version__


def test_shard_range_partitions():
    for n in (0, 1, 7, 8, 20000, 3001):
        for w in (1, 2, 3, 8):
            cuts = [dist_mod.shard_range(n, r, w) for r in range(w)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
    from oracle import retrieval_oracle as R
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        db, qs, _ = _data.retrieval_problem(D=1001, Q=40, seed=9)
        s, e = dist_mod.shard_range(len(db), rank, world)

        def local_topk(dbl, queries, k, off):          # injected checker implementation (CPU): host logic under test
            d, i = R.knn_f64(dbl, queries, k)
            return torch.from_numpy(d), torch.from_numpy(i + off)

        def merge(gd, gi):
            Rr, Q, k = gd.shape
            d = gd.permute(1, 0, 2).reshape(Q, Rr * k).numpy()
            i = gi.permute(1, 0, 2).reshape(Q, Rr * k).numpy()
            order = np.lexsort((i, d), axis=1)[:, :k]
            return torch.from_numpy(np.take_along_axis(d, order, 1)), torch.from_numpy(np.take_along_axis(i, order, 1))

        md, mi = dist_mod.retrieve_sharded(local_topk, merge, db[s:e], s, qs, 25)
        rd, ri = R.knn_f64(db, qs, 25)
        ok_r = bool(np.array_equal(mi.numpy(), ri))
        clouds = np.arange(7 * 4 * 3, dtype=np.float32).reshape(7, 4, 3)
        full, (a, b) = dist_mod.embed_sharded(lambda c: c.reshape(len(c), -1)[:, :5] * 2.0, clouds, gather=True)
        ok_e = bool(np.array_equal(full, clouds.reshape(7, -1)[:, :5] * 2.0)) and (b - a) in (3, 4)
        q.put((rank, ok_r, ok_e))
    finally:
        dist.destroy_process_group()


def test_sharded_retrieval_and_embedding_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(2)]
    for p in procs:
        p.join(60)
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] and r[2] for r in res), res


def _worker_2d(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from oracle import retrieval_oracle as R
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        db, qs, _ = _data.retrieval_problem(D=601, Q=43, dim=64, seed=13)
        grid = dist_mod.make_grid_groups(2)                      # 2 database shards x 2 query groups
        ds, qg, q_groups = grid[0], grid[1], grid[2]
        s, e = dist_mod.shard_range(len(db), ds, 2)

        def local_topk(dbl, queries, k, off):
            d, i = R.knn_f64(dbl, queries.numpy(), k)
            return torch.from_numpy(d), torch.from_numpy(i + off)

        def merge(gd, gi):
            Rr, Q, k = gd.shape
            d = gd.permute(1, 0, 2).reshape(Q, Rr * k).numpy()
            i = gi.permute(1, 0, 2).reshape(Q, Rr * k).numpy()
            order = np.lexsort((i, d), axis=1)[:, :k]
            return torch.from_numpy(np.take_along_axis(d, order, 1)), torch.from_numpy(np.take_along_axis(i, order, 1))

        md, mi = dist_mod.retrieve_sharded_2d(local_topk, merge, db[s:e], s, torch.from_numpy(qs), 25, grid)
        rd, ri = R.knn_f64(db, qs, 25)
        q.put((rank, bool(np.array_equal(mi.numpy(), ri)) and bool(np.abs(md.numpy() - rd).max() <= 1e-12), (ds, qg, q_groups)))
    finally:
        dist.destroy_process_group()


def test_sharded_retrieval_2d_world4_gloo():
    """2 database shards x 2 query groups: every rank ends with the full, shard-layout-independent result."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_2d, args=(r, 4, port, q)) for r in range(4)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(4)]
    for p in procs:
        p.join(60)
    assert sorted(r[0] for r in res) == [0, 1, 2, 3]
    assert all(r[1] for r in res), res
    assert sorted(r[2] for r in res) == [(0, 0, 2), (0, 1, 2), (1, 0, 2), (1, 1, 2)]


def test_loading_pointclouds_wire_formats(tmp_path):
    """.bin float64 clouds (utils/loading_pointclouds.py:26-65) and the evaluation pickles."""
    import pickle
    lp = importlib.import_module("epc-net_b200.utils.loading_pointclouds")
    rng = np.random.default_rng(0)
    good = rng.uniform(-1, 1, (4096, 3))
    good.astype(np.float64).tofile(tmp_path / "a.bin")
    rng.uniform(-1, 1, (100, 3)).tofile(tmp_path / "short.bin")
    pc = lp.load_pc_file("a.bin", str(tmp_path))
    assert pc.dtype == np.float64 and pc.shape == (4096, 3) and np.array_equal(pc, good)
    assert not lp.load_pc_file("short.bin", str(tmp_path)).any()                    # wrong size -> zeros (:33-38)
    assert lp.load_pc_files(["a.bin", "short.bin"], str(tmp_path)).shape == (2, 4096, 3)
    runs = [{0: {"query": "a.bin", "northing": 1.0, "easting": 2.0, 1: [0]}, 1: {"query": "short.bin", "northing": 3.0, "easting": 4.0}},
            {0: {"query": "a.bin", "northing": 1.5, "easting": 2.5, 0: [0, 1]}}]
    with open(tmp_path / "db.pickle", "wb") as f:
        pickle.dump(runs, f)
    sets = lp.get_sets_dict(str(tmp_path / "db.pickle"))
    data = lp.load_pc_data_set(sets, str(tmp_path))
    assert [d.shape for d in data] == [(2, 4096, 3), (1, 4096, 3)] and data[0].dtype == np.float32
    assert np.array_equal(data[0][0], good.astype(np.float32)) and not data[0][1].any()
    feat = rng.uniform(0, 5, (4096, 13))
    feat.tofile(tmp_path / "f.bin")
    p13 = lp.load_pc_file("f.bin", str(tmp_path), input_dim=13)
    assert p13.shape == (4096, 13) and p13[:, 3:12].min() >= 0 and p13[:, 3:12].max() <= 1 and np.array_equal(p13[:, :3], feat[:, :3])


def test_variable_store_uids_are_never_reused():
    """Engine caches key on VariableStore.uid: id() of a collected store is handed to the next one."""
    import gc
    variables = importlib.import_module("epc-net_b200.variables")
    seen = set()
    for _ in range(64):
        s = variables.VariableStore({"a": np.zeros(1, np.float32)})
        assert s.uid not in seen
        seen.add(s.uid)
        del s
        gc.collect()


def test_train_mirror_surface():
    """train.py:857-965 mirrors: the reference's names and module globals; no CUDA needed for the empty cases."""
    train = importlib.import_module("epc-net_b200.train")
    assert hasattr(train, "TRAINING_LATENT_VECTORS") and hasattr(train, "train_data")
    saved = train.train_data
    try:
        train.train_data = None
        with pytest.raises(RuntimeError):
            train.get_latent_vectors(None, {}, {0: {}})
        train.train_data = np.zeros((0, 128, 3), np.float32)
        assert train.get_latent_vectors(None, {}, {}).shape == (0,)
    finally:
        train.train_data = saved


def test_reference_style_plugin_import():
    """evaluate.py:11-13,119 verbatim: the reference appends its `models` (and `utils`) directory to sys.path and imports the
    yaml ARCH value as a TOP-LEVEL module.  With our directories in their place the unmodified statements must work."""
    import subprocess
    code = r'''
import importlib, os, sys
BASE_DIR = %r
sys.path.append(BASE_DIR)
sys.path.append(os.path.join(BASE_DIR, 'models'))
sys.path.append(os.path.join(BASE_DIR, 'utils'))
from loading_pointclouds import get_sets_dict, load_pc_file, load_pc_files
import tf_util
import loupe as lp
for arch in ("epc-net", "epc-net-l", "kd_epc-net", "kd_epc-net-l"):
    MODEL = importlib.import_module(arch)
    assert callable(MODEL.forward) and callable(MODEL.placeholder_inputs) and MODEL.ARCH == arch
assert callable(tf_util.pairwise_distance_mask) and callable(tf_util.conv1d) and callable(lp.G_VLAD) and callable(lp.NetVLAD)
print("ok")
''' % os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "epc-net_b200")
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd=tempfile.gettempdir())
    assert res.returncode == 0 and res.stdout.strip().endswith("ok"), res.stderr[-2000:]


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted (GPU box)")
@pytest.mark.parametrize("cfg", ["epc-net.yaml", "epc-net-l.yaml", "epc-net-l-d.yaml"])
def test_reference_yaml_configs_feed_the_plugin(cfg):
    """configs/*.yaml through yaml.safe_load (evaluate.py:37-82) give a params dict the plugin accepts as it is: the ARCH
    value names a plugin module, every key forward() reads is present, and the values are the ones the kernels are built for."""
    import yaml
    models = importlib.import_module("epc-net_b200.models")
    args = yaml.safe_load(open(os.path.join(REF, "configs", cfg)))
    archs = [args[k] for k in ("ARCH", "ARCH_TEACHER", "ARCH_STUDENT") if k in args]
    assert archs
    for arch in archs:
        mod = models.load(arch)
        assert mod.ARCH == arch
        for key in ("CLUSTER_SIZE", "FEATURE_OUTPUT_DIM", "KNN", "INPUT_DIM", "NUM_POINTS"):
            assert key in args, key
        assert (args["NUM_POINTS"], args["INPUT_DIM"], args["CLUSTER_SIZE"], args["FEATURE_OUTPUT_DIM"], args["KNN"]) == \
               (4096, 3, 64, 256, 20)
        # parameter validation happens before any device work: a training-mode call and a missing key fail like the reference
        with pytest.raises(NotImplementedError):
            mod.forward(np.zeros((1, 1, 4096, 3), np.float32), True, params=args)
        bad = {k: v for k, v in args.items() if k != "KNN"}
        with pytest.raises(KeyError):
            mod.forward(np.zeros((1, 1, 4096, 3), np.float32), False, params=bad)

"""Database/query recall flow of the reference's evaluate.py on the B200 path.

  get_latent_vectors  evaluate.py:351-452   clouds -> descriptors (here: batched, pipelined H2D)
  get_recall          evaluate.py:455-537   KDTree 25-NN per query -> recall@N / top-1 similarity / top-1%
  evaluate            evaluate.py:226-348   all database sets, all query sets, the m != n pair loop

Differences from the reference, by construction:
  * no ``sess``: the argument is kept (and ignored) so call sites read the same;
  * ``ops`` is a plain dict: {"MODEL": plugin module, "params": yaml dict, optional "BATCH_NUM_QUERIES",
    "POSITIVES_PER_QUERY", "NEGATIVES_PER_QUERY" (defaults 1, 0, 0 as at evaluate.py:86-90)};
  * the reference runs ONE cloud per sess.run; inference-mode batch norm makes every descriptor independent
    of its batch (utils/tf_util.py:486-489), so clouds are embedded in large chunks instead -- same values;
  * retrieval is an exact GPU k-NN (epc_retrieve_topk) instead of a CPU KD-tree -- same indices.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np
import torch

from . import _lib, variables
from . import engine as _engine
from .engine import _ptr, _stream, workspaces

# the reference keeps these as module globals (evaluate.py:130-134, 227-229)
DATABASE_SETS = []
QUERY_SETS = []
DATABASE_VECTORS = []
QUERY_VECTORS = []
NUM_NEIGHBORS = 25      # evaluate.py:465


def _arch_of(ops):
    model = ops["MODEL"]
    return getattr(model, "ARCH", None) or ops["params"]["ARCH"]


def get_latent_vectors(sess, ops, dict_to_process, data):
    """evaluate.py:351-452.  ``data``: (n, NUM_POINTS, INPUT_DIM) host array for the n entries of
    ``dict_to_process``; returns the (n, FEATURE_OUTPUT_DIM) fp32 descriptors as a host array.

    The reference's grouping into (BATCH_NUM_QUERIES x (1+P+N)) tuples and its zero "fake" clouds for the tail
    (:415-446) only pad sess.run feeds; they do not change any returned row, so rows are embedded directly.
    """
    n = len(dict_to_process.keys()) if dict_to_process is not None else len(data)
    data = np.asarray(data, dtype=np.float32)
    if data.shape[0] != n:
        raise ValueError("data has %d clouds but dict_to_process has %d entries" % (data.shape[0], n))
    params = ops["params"]
    eng = _engine.get_engine(_arch_of(ops), params, store=params.get("VARIABLES"))
    return eng.embed_host(data)


class RetrievalIndex:
    """The prepared database: stands where ``database_nbrs = KDTree(database_output)`` (evaluate.py:463) stands in the
    reference -- built once per database set, queried by every query set.  ``query(q, k)`` returns
    (dist [Q,k] float64, idx [Q,k] int64) like ``KDTree.query``, as CUDA tensors.  ``id_offset`` is added to every
    returned row id (a database shard reports global ids)."""

    def __init__(self, database_output, id_offset=0):
        lib = _lib.load()
        db = _engine.as_cuda_f32(database_output, "database_output")
        if db.dim() != 2:
            raise ValueError("RetrievalIndex expects a (D, dim) array, got %s" % (tuple(db.shape),))
        self.db = db
        self.id_offset = int(id_offset)
        D, dim = db.shape
        with torch.cuda.device(db.device):
            self.mem = torch.empty((int(lib.epc_retrieve_index_bytes(D, dim)),), dtype=torch.uint8, device=db.device)
            _lib.check(lib.epc_retrieve_index_build(_ptr(db), D, dim, _ptr(self.mem), self.mem.numel(), _stream()))

    def query(self, queries_output, k, out=None):
        """``out``: optional preallocated (dist [Q,k] float64, idx [Q,k] int64) contiguous CUDA tensors."""
        lib = _lib.load()
        db = self.db
        q = _engine.as_cuda_f32(queries_output, "queries_output")
        if q.dim() != 2 or db.shape[1] != q.shape[1]:
            raise ValueError("query expects a (Q, %d) array, got %s" % (db.shape[1], tuple(q.shape)))
        if q.device != db.device:
            q = q.to(db.device)
        D, dim = db.shape
        Q = q.shape[0]
        k = int(k)
        if out is not None:
            dist, idx = out
            if (dist.dtype != torch.float64 or idx.dtype != torch.int64 or tuple(dist.shape) != (Q, k) or tuple(idx.shape) != (Q, k)
                    or not dist.is_contiguous() or not idx.is_contiguous() or dist.device != db.device or idx.device != db.device):
                raise ValueError("out must be contiguous (float64 [Q,k], int64 [Q,k]) tensors on the database's device")
        else:
            idx = torch.empty((Q, k), dtype=torch.int64, device=db.device)
            dist = torch.empty((Q, k), dtype=torch.float64, device=db.device)
        with torch.cuda.device(db.device):
            ws = workspaces.get(lib.epc_retrieve_workspace_bytes(D, Q, dim, k))
            _lib.check(lib.epc_retrieve_topk_indexed(_ptr(db), D, _ptr(self.mem), _ptr(q), Q, dim, k, self.id_offset,
                                                     _ptr(idx), _ptr(dist), _ptr(ws), ws.numel(), _stream()))
        return dist, idx


def retrieve_topk(database_output, queries_output, k, id_offset=0):
    """Exact Euclidean k-NN on the GPU: -> (dist [Q,k] float64, idx [Q,k] int64), ascending, like
    ``KDTree(database_output).query(queries_output, k)`` (evaluate.py:463,481).  One-shot form: the database is
    prepared inside the call (see RetrievalIndex for the build-once form)."""
    lib = _lib.load()
    db = _engine.as_cuda_f32(database_output, "database_output")
    q = _engine.as_cuda_f32(queries_output, "queries_output")
    if db.dim() != 2 or q.dim() != 2 or db.shape[1] != q.shape[1]:
        raise ValueError("retrieve_topk expects (D, dim) and (Q, dim) arrays, got %s and %s" % (tuple(db.shape), tuple(q.shape)))
    if q.device != db.device:
        q = q.to(db.device)
    D, dim = db.shape
    Q = q.shape[0]
    k = int(k)
    idx = torch.empty((Q, k), dtype=torch.int64, device=db.device)
    dist = torch.empty((Q, k), dtype=torch.float64, device=db.device)
    with torch.cuda.device(db.device):
        ws = workspaces.get(lib.epc_retrieve_workspace_bytes(D, Q, dim, k))
        _lib.check(lib.epc_retrieve_topk(_ptr(db), D, _ptr(q), Q, dim, k, int(id_offset), _ptr(idx), _ptr(dist), _ptr(ws),
                                         ws.numel(), _stream()))
    return dist, idx


def query_radius(database_coords, query_coords, r):
    """``KDTree(database_coords).query_radius(query_coords, r)`` (generating_queries/generate_test_sets.py:70-104: r = 25 m
    on (northing, easting); generate_training_tuples_baseline.py:52-62: r = 10 / 50) as one exact GPU search in float64.
    Returns a list with one ascending int64 index array per query (sklearn returns the same sets in tree order)."""
    _engine._require_cuda()
    lib = _lib.load()
    db = torch.as_tensor(np.ascontiguousarray(database_coords, dtype=np.float64)).cuda()
    q = torch.as_tensor(np.ascontiguousarray(query_coords, dtype=np.float64)).cuda()
    if db.dim() != 2 or q.dim() != 2 or db.shape[1] != q.shape[1]:
        raise ValueError("query_radius expects (D, dim) and (Q, dim) arrays")
    D, dim = db.shape
    Q = q.shape[0]
    if Q == 0:
        return []
    counts = torch.zeros((Q,), dtype=torch.int32, device=db.device)
    with torch.cuda.device(db.device):
        _lib.check(lib.epc_radius_count(_ptr(db), D, _ptr(q), Q, dim, float(r), _ptr(counts), _stream()))
        offsets = torch.cumsum(counts.to(torch.int64), 0) - counts.to(torch.int64)
        total = int(counts.sum().item())
        indices = torch.empty((max(total, 1),), dtype=torch.int32, device=db.device)
        _lib.check(lib.epc_radius_fill(_ptr(db), D, _ptr(q), Q, dim, float(r), _ptr(offsets), _ptr(indices), _stream()))
    ind = indices[:total].cpu().numpy().astype(np.int64)
    off = offsets.cpu().numpy()
    cnt = counts.cpu().numpy()
    return [ind[off[i]:off[i] + cnt[i]] for i in range(Q)]


def get_random_hard_negatives(query_vec, random_negs, num_to_take, training_latent_vectors):
    """train.py:857-869: the ``num_to_take`` candidates of ``random_negs`` whose cached descriptors are
    nearest to ``query_vec`` -- the reference builds a KDTree over ``TRAINING_LATENT_VECTORS[random_negs]`` per query;
    here it is one exact GPU k-NN (same indices).  ``training_latent_vectors`` replaces the reference's global."""
    random_negs = np.asarray(random_negs)
    latent_vecs = np.asarray(training_latent_vectors, dtype=np.float32)[random_negs]
    _, indices = retrieve_topk(latent_vecs, np.asarray(query_vec, dtype=np.float32)[None], int(num_to_take))
    hard_negs = np.squeeze(random_negs[indices[0].cpu().numpy()])
    return hard_negs.tolist()


def recall_from_neighbors(indices, valid, true_neighbors_list, queries_output, database_output, num_db,
                          num_neighbors=NUM_NEIGHBORS):
    """The bookkeeping of evaluate.py:466-530 given the retrieved ``indices`` (one row per evaluated query)."""
    recall = [0] * num_neighbors
    top1_similarity_score = []
    one_percent_retrieved = 0
    threshold = max(int(round(num_db / 100.0)), 1)                              # :470
    for row, i in enumerate(valid):
        true_neighbors = true_neighbors_list[row]
        ind = indices[row]
        tset = set(int(t) for t in true_neighbors)
        for j in range(len(ind)):                                               # :513
            if int(ind[j]) in tset:
                if j == 0:
                    top1_similarity_score.append(np.dot(queries_output[i], database_output[ind[j]]))   # :516
                recall[j] += 1
                break
        if len(set(int(v) for v in ind[0:threshold]).intersection(tset)) > 0:    # :526
            one_percent_retrieved += 1
    num_evaluated = len(valid)
    one_percent_recall = (one_percent_retrieved / float(num_evaluated)) * 100   # :529
    recall = (np.cumsum(recall) / float(num_evaluated)) * 100                   # :530
    return recall, top1_similarity_score, one_percent_recall


_INDEX_CACHE = {}       # m -> (DATABASE_VECTORS[m] object, RetrievalIndex): one "KDTree" per database set, not per (m, n) pair


def _database_index(m):
    vec = DATABASE_VECTORS[m]
    hit = _INDEX_CACHE.get(m)
    if hit is None or hit[0] is not vec:
        for key in [key for key, v in _INDEX_CACHE.items() if not any(v[0] is d for d in DATABASE_VECTORS)]:
            del _INDEX_CACHE[key]                                               # sets of an earlier evaluation
        hit = (vec, RetrievalIndex(np.asarray(vec)))
        _INDEX_CACHE[m] = hit
    return hit[1]


def get_recall(sess, ops, m, n, fout=None):
    """evaluate.py:455-537 on the module globals DATABASE_VECTORS / QUERY_VECTORS / QUERY_SETS.
    Returns (recall, top1_similarity_score, one_percent_recall, for_plot)."""
    database_output = np.asarray(DATABASE_VECTORS[m])
    queries_output = np.asarray(QUERY_VECTORS[n])
    valid, truth = [], []
    for i in range(len(queries_output)):
        true_neighbors = QUERY_SETS[n][i][m]                                    # :477
        if len(true_neighbors) == 0:                                            # :478
            continue
        valid.append(i)
        truth.append(true_neighbors)
    if not valid:
        raise ZeroDivisionError("no query of set %d has a true neighbour in set %d" % (n, m))   # :529 divides by 0
    k = min(NUM_NEIGHBORS, len(database_output))
    _, idx = _database_index(m).query(queries_output[valid], k)
    idx = idx.cpu().numpy()
    recall, sim, opr = recall_from_neighbors(idx, valid, truth, queries_output, database_output, len(database_output))
    for_plot = []
    for row, i in enumerate(valid):                                             # :507-524
        q = QUERY_SETS[n][i]
        for_plot.append(q.get("easting"))
        for_plot.append(q.get("northing"))
        tset = set(int(t) for t in truth[row])
        hit = [j for j in range(idx.shape[1]) if int(idx[row, j]) in tset]
        for_plot.append(hit[0] if hit else 25)
        if fout is not None:
            fout.write("%s|%s|%s\n" % (q.get("query", ""), " ".join(str(int(v)) for v in idx[row]),
                                       " ".join(str(int(t)) for t in truth[row])))
    return recall, sim, opr, for_plot


def evaluate(ops, eval_database_set, eval_query_set, database_sets, query_sets, output_file=None):
    """evaluate.py:226-348 without the TF graph/session plumbing.

    eval_database_set[i] / eval_query_set[j]: (n, NUM_POINTS, 3) host arrays (what load_pc_data_set returns,
    evaluate.py:177-195); database_sets / query_sets: the evaluation pickles' dict lists.
    Returns (ave_recall[25], average_similarity, ave_one_percent_recall).
    """
    global DATABASE_SETS, QUERY_SETS, DATABASE_VECTORS, QUERY_VECTORS
    DATABASE_SETS, QUERY_SETS = database_sets, query_sets
    DATABASE_VECTORS, QUERY_VECTORS = [], []
    recall = np.zeros(NUM_NEIGHBORS)
    count = 0
    similarity = []
    one_percent_recall = []
    for i in range(len(DATABASE_SETS)):                                         # :297-299
        DATABASE_VECTORS.append(get_latent_vectors(None, ops, DATABASE_SETS[i], eval_database_set[i]))
    for j in range(len(QUERY_SETS)):                                            # :301-303
        QUERY_VECTORS.append(get_latent_vectors(None, ops, QUERY_SETS[j], eval_query_set[j]))
    for m in range(len(QUERY_SETS)):                                            # :305-317
        for n in range(len(QUERY_SETS)):
            if m == n:
                continue
            pair_recall, pair_similarity, pair_opr, _ = get_recall(None, ops, m, n, fout=None)
            recall += np.array(pair_recall)
            count += 1
            one_percent_recall.append(pair_opr)
            similarity.extend(pair_similarity)
    ave_recall = recall / count                                                 # :320
    average_similarity = np.mean(similarity)                                    # :326
    ave_one_percent_recall = np.mean(one_percent_recall)                        # :330
    if output_file:
        os.makedirs(os.path.dirname(os.path.abspath(output_file)), exist_ok=True)
        with open(output_file, "a") as output:                                  # :336-348
            output.write(str(ops["params"].get("ARCH", _arch_of(ops))))
            output.write("\n\nAverage Recall @N:\n")
            output.write(str(ave_recall))
            output.write("\n\nAverage Similarity:\n")
            output.write(str(average_similarity))
            output.write("\n\nAverage Top 1% Recall:\n")
            output.write(str(ave_one_percent_recall))
            output.write("\n\n")
    return ave_recall, average_similarity, ave_one_percent_recall

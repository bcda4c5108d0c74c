"""CPU ORACLE (test infrastructure, NOT a product path): retrieval / recall of evaluate.py.

The reference ranks with ``sklearn.neighbors.KDTree(database_output).query(q[None], k=25)``
(evaluate.py:463,481): Euclidean distance evaluated in float64 on the fp32 descriptors, ascending.
``knn_f64`` restates that as brute force (a KD-tree is exact, so the answer is the same up to exact
float64 ties); ``tests/golden/retrieval_kdtree.npz`` pins it against the real sklearn KDTree.
"""
from __future__ import annotations

import numpy as np


def knn_f64(db, q, k):
    """-> (dist [Q,k] float64 ascending, idx [Q,k] int64); ties -> lower index first."""
    db = np.asarray(db, dtype=np.float64)
    q = np.asarray(q, dtype=np.float64)
    k = min(k, db.shape[0])
    dist = np.empty((q.shape[0], k), np.float64)
    idx = np.empty((q.shape[0], k), np.int64)
    for s in range(0, q.shape[0], 256):
        qq = q[s:s + 256]
        d2 = ((qq[:, None, :] - db[None, :, :]) ** 2).sum(-1) if db.shape[0] * qq.shape[0] * db.shape[1] < 5e7 else \
            np.stack([((db - v) ** 2).sum(-1) for v in qq], 0)
        order = np.argsort(d2, axis=1, kind="stable")[:, :k]
        idx[s:s + 256] = order
        dist[s:s + 256] = np.sqrt(np.take_along_axis(d2, order, 1))
    return dist, idx


def get_recall(database_output, queries_output, query_set, m, num_neighbors=25, knn_fn=None):
    """evaluate.get_recall -- evaluate.py:455-537 (print_log/for_plot bookkeeping omitted).

    query_set: QUERY_SETS[n] = {i: {m: [true neighbour indices in database m], ...}}.
    Returns (recall[num_neighbors] cumulative %, top1_similarity list, one_percent_recall %).
    """
    database_output = np.asarray(database_output)
    queries_output = np.asarray(queries_output)
    knn_fn = knn_fn or knn_f64
    recall = [0] * num_neighbors                                               # :466
    top1_similarity_score = []
    one_percent_retrieved = 0
    threshold = max(int(round(len(database_output) / 100.0)), 1)               # :470
    num_evaluated = 0
    for i in range(len(queries_output)):                                       # :476
        true_neighbors = query_set[i][m]                                       # :477
        if len(true_neighbors) == 0:                                           # :478
            continue
        num_evaluated += 1
        _, indices = knn_fn(database_output, queries_output[i][None], num_neighbors)   # :481
        for j in range(len(indices[0])):                                       # :513
            if indices[0][j] in true_neighbors:
                if j == 0:
                    similarity = np.dot(queries_output[i], database_output[indices[0][j]])      # :516
                    top1_similarity_score.append(similarity)
                recall[j] += 1
                break
        if len(set(indices[0][0:threshold]).intersection(set(true_neighbors))) > 0:               # :526
            one_percent_retrieved += 1
    one_percent_recall = (one_percent_retrieved / float(num_evaluated)) * 100  # :529
    recall = (np.cumsum(recall) / float(num_evaluated)) * 100                  # :530
    return recall, top1_similarity_score, one_percent_recall


def evaluate_pairs(database_vectors, query_vectors, query_sets, num_neighbors=25, knn_fn=None):
    """The m != n pair loop and averaging of evaluate.py:305-334."""
    recall = np.zeros(num_neighbors)
    count = 0
    similarity = []
    one_percent_recall = []
    for m in range(len(query_sets)):
        for n in range(len(query_sets)):
            if m == n:
                continue
            pr, ps, po = get_recall(database_vectors[m], query_vectors[n], query_sets[n], m, num_neighbors, knn_fn)
            recall += np.array(pr)
            count += 1
            one_percent_recall.append(po)
            similarity.extend(ps)
    return recall / count, float(np.mean(similarity)) if similarity else float("nan"), float(np.mean(one_percent_recall))


def kdtree_knn(db, q, k):
    """The reference's exact call (needs scikit-learn)."""
    from sklearn.neighbors import KDTree
    tree = KDTree(np.asarray(db))                                              # evaluate.py:463
    d, i = tree.query(np.asarray(q), k=k)                                      # evaluate.py:481
    return d, i

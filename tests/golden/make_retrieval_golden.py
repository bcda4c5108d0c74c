"""Mint tests/golden/retrieval_kdtree.npz with the reference's own retrieval library call:
sklearn.neighbors.KDTree(db).query(q[None], k=25) per query (evaluate.py:463,481)."""
import os
import sys

import numpy as np
from sklearn.neighbors import KDTree
import sklearn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE)))
import _data  # noqa: E402

db, q, src = _data.retrieval_problem(D=4000, Q=500, seed=7)
tree = KDTree(db)
dist = np.empty((len(q), 25)); idx = np.empty((len(q), 25), np.int64)
for i in range(len(q)):                       # one query at a time, as the reference does
    d, j = tree.query(np.array([q[i]]), k=25)
    dist[i], idx[i] = d[0], j[0]
np.savez_compressed(os.path.join(HERE, "retrieval_kdtree.npz"), D=4000, Q=500, seed=7, dist=dist, idx=idx,
                    sklearn_version=sklearn.__version__)
print("wrote retrieval_kdtree.npz; recall@1 =", (idx[:, 0] == src).mean())

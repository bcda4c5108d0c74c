"""Training-side callers of the embedding/retrieval path: the two helpers ``train.py`` runs every 1400
iterations to refresh its hard-negative cache.  Training itself (losses, optimiser) is out of scope.

  get_latent_vectors(sess, ops, dict_to_process)          train.py:871-965   embed the whole training set
  get_random_hard_negatives(query_vec, random_negs, n)    train.py:857-869   n nearest cached descriptors among the candidates

Like the reference, both work on module globals: ``train_data`` (the (n, NUM_POINTS, INPUT_DIM) array train.py builds at
start-up) and ``TRAINING_LATENT_VECTORS`` (the cache get_latent_vectors' caller fills, train.py:196, 291).
"""
from __future__ import annotations

import numpy as np

from . import evaluate as _evaluate

train_data = None                   # train.py: global array of all training submaps
TRAINING_LATENT_VECTORS = []        # train.py:57


def get_latent_vectors(sess, ops, dict_to_process):
    """train.py:871-965: descriptors of ``train_data[0 : len(dict_to_process)]`` in dict order.

    The reference packs tuples of BATCH_NUM_QUERIES*(1+P+N+1) clouds into the four placeholders and pads the tail with
    zero clouds; with inference batch norm that grouping changes no row, so the rows are embedded directly.  Its result
    shapes are kept: (n, D) for n >= 2, a flat (D,) vector for n == 1 (the tail loop's ``np.squeeze``), an empty array for
    n == 0."""
    if train_data is None:
        raise RuntimeError("set train.train_data (the (n, NUM_POINTS, INPUT_DIM) training array) first")
    n = len(dict_to_process.keys())
    if n == 0:
        return np.array([])
    out = _evaluate.get_latent_vectors(sess, ops, dict_to_process, np.asarray(train_data)[:n])
    return out[0] if n == 1 else out


def get_random_hard_negatives(query_vec, random_negs, num_to_take):
    """train.py:857-869 on the module global TRAINING_LATENT_VECTORS."""
    return _evaluate.get_random_hard_negatives(query_vec, random_negs, num_to_take, TRAINING_LATENT_VECTORS)

"""Dataset wire formats on either side of the hot path, with the names of the reference's
utils/loading_pointclouds.py.  Host-side numpy only: these feed ``evaluate.get_latent_vectors`` (which pins and streams
the clouds to the GPU), they do not compute.

  load_pc_file / load_pc_files   utils/loading_pointclouds.py:26-65   ``.bin`` = raw little-endian float64, 4096 x 3 (or x 13)
  get_sets_dict / get_queries_dict   :11-24                           pickles written by generating_queries/*.py
  load_pc_data / load_pc_data_set    evaluate.py:154-195              a whole evaluation run as one (n, 4096, 3) fp32 array

The 13-column (handcrafted-feature) variant is parsed like the reference but the embedding path itself is xyz-only
(INPUT_DIM = 3: conv1 of every shipped checkpoint is [1,3,64]).
"""
from __future__ import annotations

import os
import pickle

import numpy as np

NUM_POINTS = 4096       # utils/loading_pointclouds.py:32,41


def get_queries_dict(filename):
    """:11-16 -- {key: {'query': file, 'positives': [...], 'negatives': [...]}} (training tuples)."""
    with open(filename, "rb") as handle:
        return pickle.load(handle)


def get_sets_dict(filename):
    """:18-24 -- [ {key: {'query': file, 'northing': v, 'easting': v, <run m>: [true-neighbour keys], ...}}, ... ]
    (one dict per run; written by generating_queries/generate_test_sets.py:77-109)."""
    with open(filename, "rb") as handle:
        return pickle.load(handle)


def _read_cloud(path, columns):
    """Raw float64 records -> (NUM_POINTS, columns), or None when the file does not hold exactly that many values."""
    flat = np.fromfile(path, dtype="<f8")
    return flat.reshape(NUM_POINTS, columns) if flat.size == NUM_POINTS * columns else None


def load_pc_file(filename, dataset_folder, input_dim=3):
    """:26-53.  A file of the wrong size yields an all-zero cloud (the reference prints an error and does the same).
    With the 13-column layout, columns 3..11 are min-max scaled per cloud, NaN -> 0 and inf -> 1 (:47-51)."""
    columns = 3 if input_dim == 3 else 13
    cloud = _read_cloud(os.path.join(dataset_folder, filename), columns)
    if cloud is None:
        return np.zeros((NUM_POINTS, columns))
    if columns == 13:
        lo, hi = cloud.min(axis=0), cloud.max(axis=0)
        with np.errstate(divide="ignore", invalid="ignore"):
            cloud[:, 3:12] = ((cloud - lo) / (hi - lo))[:, 3:12]
        np.nan_to_num(cloud, copy=False, nan=0.0, posinf=1.0, neginf=1.0)
    return cloud


def load_pc_files(filenames, dataset_folder, input_dim=3):
    """:56-65 -- stacks the clouds that have 4096 points."""
    clouds = (load_pc_file(name, dataset_folder, input_dim) for name in filenames)
    return np.array([c for c in clouds if c.shape[0] == NUM_POINTS])


def load_pc_data(data, dataset_folder, input_dim=3, out=None):
    """evaluate.py:154-184 -- every cloud of one run (``data[i]['query']``, i = 0..len-1) as (n, 4096, input_dim) fp32.
    ``out`` may be a preallocated (pinned) buffer of that shape."""
    n = len(data.keys())
    if out is None:
        out = np.empty((n, NUM_POINTS, input_dim), np.float32)
    for i in range(n):
        out[i] = load_pc_file(data[i]["query"], dataset_folder, input_dim)       # float64 -> fp32, evaluate.py:169
    return out[:n]


def load_pc_data_set(data_set, dataset_folder, input_dim=3):
    """evaluate.py:186-195."""
    return [load_pc_data(d, dataset_folder, input_dim) for d in data_set]

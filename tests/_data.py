"""Seeded synthetic inputs shared by the tests, the golden generator and bench.py."""
from __future__ import annotations

import numpy as np

N_POINTS = 4096


def cloud(kind: str, seed: int, n: int = N_POINTS) -> np.ndarray:
    """One (n,3) fp32 cloud.  Kinds: uniform | quantised (1/64 grid => mass ties) | duplicated |
    planar | zeros (the all-zero "fake" cloud of evaluate.py:425-430) | clustered."""
    rng = np.random.default_rng(seed)
    if kind == "uniform":
        p = rng.uniform(-1.0, 1.0, (n, 3))
    elif kind == "quantised":
        p = np.round(rng.uniform(-1.0, 1.0, (n, 3)) * 64.0) / 64.0
    elif kind == "coarse":
        p = np.round(rng.uniform(-1.0, 1.0, (n, 3)) * 6.0) / 6.0      # ~2200 distinct sites: heavy ties
    elif kind == "duplicated":
        base = rng.uniform(-1.0, 1.0, ((n + 2) // 3, 3))           # 3 copies: 20 = 6*3+2 splits a tie group
        p = np.concatenate([base] * 3, 0)[:n][rng.permutation(n)]
    elif kind == "planar":
        p = rng.uniform(-1.0, 1.0, (n, 3))
        p[:, 2] = 0.25
    elif kind == "zeros":
        p = np.zeros((n, 3))
    elif kind == "clustered":
        c = rng.uniform(-0.8, 0.8, (16, 3))
        p = c[rng.integers(0, 16, n)] + rng.normal(0, 0.05, (n, 3))
    else:
        raise ValueError(kind)
    return np.ascontiguousarray(p, dtype=np.float32)


GOLDEN_KINDS = ["uniform"] * 12 + ["clustered", "planar", "quantised", "coarse", "duplicated", "zeros"]


def golden_batch(seed: int = 1000, n: int = N_POINTS) -> np.ndarray:
    """The 18-cloud training-tuple-sized batch (train.py:238-255) fed to the shipped graphs."""
    return np.stack([cloud(k, seed + i, n) for i, k in enumerate(GOLDEN_KINDS)], 0)


def default_params(arch: str) -> dict:
    """The keys of configs/*.yaml that the hot path reads."""
    p = {"ARCH": arch, "NUM_POINTS": 4096, "INPUT_DIM": 3, "CLUSTER_SIZE": 64, "FEATURE_OUTPUT_DIM": 256,
         "KNN": 20}
    if "epc-net-l" not in arch:
        p["GROUPS"] = 4
    return p


def _unit(a):
    return (a / np.linalg.norm(a, axis=1, keepdims=True)).astype(np.float32)


def retrieval_problem(D=20000, Q=3000, dim=256, noise=0.05, seed=7):
    """unit-norm db, queries = perturbed db rows, true neighbour = the source row."""
    rng = np.random.default_rng(seed)
    db = _unit(rng.standard_normal((D, dim)))
    src = np.random.default_rng(seed + 1).permutation(D)[:Q] if Q <= D else np.random.default_rng(seed + 1).integers(0, D, Q)
    q = _unit(db[src] + noise * np.random.default_rng(seed + 2).standard_normal((Q, dim)).astype(np.float32))
    return db, q, src.astype(np.int64)


def retrieval_sets(n_sets=3, per_set=(400, 380, 420), dim=256, seed=21):
    """A small multi-run problem shaped like the Oxford evaluation pickles (generate_test_sets.py:77-109):
    DATABASE_SETS[m] vectors, QUERY_SETS[n][i][m] = true neighbour indices of query i of run n in run m
    (possibly empty).  Descriptors are synthetic: 'places' on a line, one noisy unit vector per visit."""
    rng = np.random.default_rng(seed)
    n_places = 500
    place_vec = _unit(rng.standard_normal((n_places, dim)))
    db_vecs, q_vecs, places = [], [], []
    for s in range(n_sets):
        pl = np.sort(rng.choice(n_places, per_set[s], replace=False))
        places.append(pl)
        v = _unit(place_vec[pl] + 0.06 * rng.standard_normal((len(pl), dim)).astype(np.float32))
        db_vecs.append(v)
        q_vecs.append(v)            # evaluate.py embeds the same submaps as database and as query
    query_sets = []
    for n in range(n_sets):
        qs = {}
        for i, p in enumerate(places[n]):
            e = {"query": "run%d/%d.bin" % (n, i), "northing": float(p), "easting": 0.0}
            for m in range(n_sets):
                e[m] = [int(j) for j in np.nonzero(np.abs(places[m].astype(int) - int(p)) <= 2)[0]] if m != n else []
            qs[i] = e
        query_sets.append(qs)
    return db_vecs, q_vecs, query_sets

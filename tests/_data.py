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

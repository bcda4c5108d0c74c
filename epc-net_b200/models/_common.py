"""Shared body of the four model plugins."""
from __future__ import annotations

import numpy as np
import torch

from .. import engine as _engine
from .. import variables


def placeholder_inputs(batch_num_queries, num_pointclouds_per_query, num_point, input_dim=13):
    """models/epc-net.py:24-26 -- the shape contract (Bq, P, N, dim) fp32; here an uninitialised CUDA
    tensor the caller fills (there is no graph/feed_dict)."""
    _engine._require_cuda()
    return torch.empty((batch_num_queries, num_pointclouds_per_query, num_point, input_dim), dtype=torch.float32,
                       device="cuda")


def forward(arch, point_cloud, is_training, bn_decay=None, params=None):
    if params is None:
        raise ValueError("params (the yaml dict) is required: CLUSTER_SIZE, FEATURE_OUTPUT_DIM, KNN, INPUT_DIM[, GROUPS]")
    if isinstance(is_training, torch.Tensor):
        is_training = bool(is_training.item())
    if is_training:
        raise NotImplementedError("epc-net_b200 implements the inference path only (is_training=False): batch-norm "
                                  "uses the stored moving statistics (utils/tf_util.py:486-489)")
    for key in ("CLUSTER_SIZE", "FEATURE_OUTPUT_DIM", "KNN", "INPUT_DIM"):
        if key not in params:
            raise KeyError(key)                  # same failure mode as params["..."] in the reference
    _engine._require_cuda()
    if isinstance(point_cloud, np.ndarray):
        point_cloud = torch.from_numpy(np.ascontiguousarray(point_cloud, dtype=np.float32)).cuda()
    if point_cloud.dim() != 4:
        raise ValueError("point_cloud must be (batch_num_queries, num_pointclouds_per_query, num_points, input_dim)")
    Bq, P, N, dim = point_cloud.shape
    if dim != params["INPUT_DIM"]:
        raise ValueError("point_cloud last dim %d != INPUT_DIM %d" % (dim, params["INPUT_DIM"]))
    out_dim = params["FEATURE_OUTPUT_DIM"]
    eng = _engine.get_engine(arch, params, store=params.get("VARIABLES"))
    flat = point_cloud.reshape(Bq * P, N, dim)                    # models/epc-net.py:41
    if Bq * P == 0:
        out = torch.empty((Bq, P, out_dim), dtype=torch.float32, device=point_cloud.device)
        if arch.startswith("kd_"):
            return torch.empty((0, 1024), dtype=torch.float32, device=point_cloud.device), out
        return out
    if arch.startswith("kd_"):
        out, feat = eng.embed(flat, want_feat=True)
        return feat, out.reshape(Bq, P, out_dim)                  # models/kd_epc-net.py:157-158
    return eng.embed(flat).reshape(Bq, P, out_dim)                # models/epc-net.py:155 "last_output"

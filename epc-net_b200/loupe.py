"""Pooling interface of the reference's loupe.py (PoolingBaseModel / NetVLAD / G_VLAD).

Same constructor arguments and ``forward(reshaped_input)`` contract as loupe.py:34-59,103-214,216-333:
``reshaped_input`` is ``[batch*max_samples, feature_size]`` (rows already L2-normalised by the caller,
models/epc-net.py:147-148) and the result is ``[batch, output_dim]`` (NOT normalised; the model does that,
models/epc-net.py:153).  Variables are read from the variable store under the current
``variables.variable_scope`` with the reference's names (cluster_weights, cluster_bn/*, cluster_weights2,
hidden1_weights, bn/*, gating_weights, gating_bn/*).  Compute: epc_vlad_forward in libepc_b200.
"""
from __future__ import annotations

import torch

if __package__:
    from . import engine as _engine
    from . import variables
else:
    # imported the reference's way -- `import loupe as lp` with the package directory on sys.path (models/epc-net.py:13-16)
    import importlib as _il
    import os as _os
    import sys as _sys
    _root = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
    if _root not in _sys.path:
        _sys.path.insert(0, _root)
    _engine = _il.import_module("epc-net_b200.engine")
    variables = _il.import_module("epc-net_b200.variables")


class PoolingBaseModel(object):
    """Inherit from this class when implementing new models (loupe.py:34-59)."""

    def __init__(self, feature_size, max_samples, cluster_size, output_dim, gating=True, add_batch_norm=True,
                 is_training=True):
        self.feature_size = feature_size
        self.max_samples = max_samples
        self.output_dim = output_dim
        self.is_training = is_training
        self.gating = gating
        self.add_batch_norm = add_batch_norm
        self.cluster_size = cluster_size

    def forward(self, reshaped_input):
        raise NotImplementedError("Models should implement the forward pass.")

    # shared by NetVLAD and G_VLAD
    def _forward(self, reshaped_input, pooling, groups):
        if isinstance(self.is_training, torch.Tensor):
            self.is_training = bool(self.is_training.item())
        if self.is_training:
            raise NotImplementedError("inference only: construct the pooling layer with is_training=False")
        if not self.add_batch_norm:
            raise NotImplementedError("add_batch_norm=False (cluster_biases / gating_biases) is not used by any "
                                      "EPC-Net config and is not implemented")
        if self.feature_size != 1024:
            raise ValueError("feature_size must be 1024 (conv5 width of every EPC-Net variant)")
        scope = variables.current_scope("query_triplets/VLAD")
        params = {"CLUSTER_SIZE": self.cluster_size, "FEATURE_OUTPUT_DIM": self.output_dim, "GROUPS": groups,
                  "KNN": 20, "INPUT_DIM": 3}
        store = getattr(self, "variables", None) or variables.default_store()
        key = ("loupe", store.uid, store.version, scope, pooling, self.gating, self.cluster_size, self.output_dim, groups,
               torch.cuda.current_device() if torch.cuda.is_available() else -1)
        eng = _ENGINES.pop(key, None)
        if eng is None:
            eng = _engine.Engine("epc-net", store, scope, params, pooling=pooling, gating=self.gating, head_only=True,
                                 vlad_prefix=scope + "/")
            while len(_ENGINES) >= 16:                        # least recently used first
                _ENGINES.pop(next(iter(_ENGINES)))
        _ENGINES[key] = eng
        return eng.vlad(reshaped_input, self.max_samples)


_ENGINES = {}


class NetVLAD(PoolingBaseModel):
    """Creates a NetVLAD class (loupe.py:103-214): hidden1_weights is [cluster_size*feature_size, output_dim]."""

    def __init__(self, feature_size, max_samples, cluster_size, output_dim, gating=True, add_batch_norm=True,
                 is_training=True):
        super().__init__(feature_size=feature_size, max_samples=max_samples, cluster_size=cluster_size,
                         output_dim=output_dim, gating=gating, add_batch_norm=add_batch_norm, is_training=is_training)

    def forward(self, reshaped_input):
        return self._forward(reshaped_input, "NetVLAD", 1)


class G_VLAD(PoolingBaseModel):
    """Creates a G_VLAD class (loupe.py:216-333): the flattened VLAD is split into ``groups`` slices that share
    one [cluster_size*feature_size/groups, output_dim] FC, batch-normalised and summed."""

    def __init__(self, feature_size, max_samples, cluster_size, output_dim, groups=4, gating=True,
                 add_batch_norm=True, is_training=True):
        super().__init__(feature_size=feature_size, max_samples=max_samples, cluster_size=cluster_size,
                         output_dim=output_dim, gating=gating, add_batch_norm=add_batch_norm, is_training=is_training)
        self.groups = groups

    def forward(self, reshaped_input):
        return self._forward(reshaped_input, "G_VLAD", self.groups)

"""EPC-Net-L: kNN graph -> 2 ProxyConv blocks -> concat 128 -> conv5 1024 -> global max-pool -> FC 256 + BN + ReLU -> L2.  Replaces models/epc-net-l.py:24-26 and :29-102.

Same plugin surface as the reference module of the same name: ``placeholder_inputs`` and
``forward(point_cloud, is_training, bn_decay=None, params=None)``.  ``point_cloud`` is a CUDA fp32 tensor
(or a numpy array, copied to the current device) of shape (Bq, P, N, INPUT_DIM); weights are looked up by their
TensorFlow names in ``params["VARIABLES"]`` or the default ``variables`` store under the current
``variables.variable_scope``.  The whole forward is one call into libepc_b200 (epc_embed).
"""
from . import _common

ARCH = "epc-net-l"
placeholder_inputs = _common.placeholder_inputs


def forward(point_cloud, is_training, bn_decay=None, params=None):
    return _common.forward(ARCH, point_cloud, is_training, bn_decay=bn_decay, params=params)

"""KD student (EPC-Net-L-D): EPC-Net-L under scope BACKBONE, also returning the per-point features.  Replaces models/kd_epc-net-l.py:29-102 (scope at :44, return value at :102).

Same plugin surface as the reference module of the same name: ``placeholder_inputs`` and
``forward(point_cloud, is_training, bn_decay=None, params=None)``.  ``point_cloud`` is a CUDA fp32 tensor
(or a numpy array, copied to the current device) of shape (Bq, P, N, INPUT_DIM); weights are looked up by their
TensorFlow names in ``params["VARIABLES"]`` or the default ``variables`` store under the current
``variables.variable_scope``.  The whole forward is one call into libepc_b200 (epc_embed).
"""
if __package__:
    from . import _common
else:
    # imported the reference's way -- sys.path.append(<models dir>); importlib.import_module(args["ARCH"]) (evaluate.py:11-13,119)
    import importlib as _il
    import os as _os
    import sys as _sys
    _root = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
    if _root not in _sys.path:
        _sys.path.insert(0, _root)
    _common = _il.import_module("epc-net_b200.models._common")

ARCH = "kd_epc-net-l"
placeholder_inputs = _common.placeholder_inputs


def forward(point_cloud, is_training, bn_decay=None, params=None):
    return _common.forward(ARCH, point_cloud, is_training, bn_decay=bn_decay, params=params)

"""Operator wrappers with the names and argument meaning of the reference's utils/tf_util.py, bound to
libepc_b200 (no TensorFlow, no CPU path).  Inputs/outputs are CUDA fp32 torch tensors.

Hot functions of the path: pairwise_distance_mask (:647-666), conv1d (:52-107), fully_connected (:310-346),
max_pool2d (:349-372); pairwise_distance (:577-596) and knn (:599-610) are exports of the same kNN kernel.
The model plugins do NOT call these one by one -- ``forward`` is a single fused call -- they exist so that
code written against the reference's operator surface keeps working and so each operator can be tested.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

if __package__:
    from .. import _lib, variables
    from ..engine import _ptr, _require_cuda, _stream, as_cuda_f32, workspaces
else:
    # imported the reference's way -- `import tf_util` with the utils directory on sys.path (models/epc-net.py:11-14)
    import importlib as _il
    import os as _os
    import sys as _sys
    _root = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
    if _root not in _sys.path:
        _sys.path.insert(0, _root)
    _lib = _il.import_module("epc-net_b200._lib")
    variables = _il.import_module("epc-net_b200.variables")
    _eng = _il.import_module("epc-net_b200.engine")
    _ptr, _require_cuda, _stream, workspaces = _eng._ptr, _eng._require_cuda, _eng._stream, _eng.workspaces
    as_cuda_f32 = _eng.as_cuda_f32

_fp = ctypes.POINTER(ctypes.c_float)


def _check_pc(pc):
    _require_cuda()
    if not (isinstance(pc, torch.Tensor) and pc.is_cuda and pc.dtype == torch.float32 and pc.dim() == 3 and pc.shape[2] == 3):
        raise ValueError("expected a CUDA fp32 tensor of shape (B, N, 3)")
    return pc.contiguous()


def pairwise_distance_mask(pc, k=20, arith="muladd"):
    """utils/tf_util.py:647-666: (B,N,3) -> (B,N,N) fp32 0/1, mask_ij = (a_ij >= 20th largest a_i.).
    As in the reference the threshold rank is the literal 20; ``k`` is accepted and ignored."""
    pc = _check_pc(pc)
    lib = _lib.load()
    B, N, _ = pc.shape
    mask = torch.empty((B, N, N), dtype=torch.float32, device=pc.device)
    with torch.cuda.device(pc.device):
        ws = workspaces.get(lib.epc_knn_workspace_bytes(B, N))
        _lib.check(lib.epc_knn_dense(_ptr(pc), B, N, _lib.KNN_ARITH[arith], _ptr(mask), None, _ptr(ws), ws.numel(), _stream()))
    return mask


def pairwise_distance(point_cloud, arith="muladd"):
    """utils/tf_util.py:577-596: (B,N,3) -> (B,N,N) squared distances in the expanded form."""
    pc = _check_pc(point_cloud)
    lib = _lib.load()
    B, N, _ = pc.shape
    dist = torch.empty((B, N, N), dtype=torch.float32, device=pc.device)
    with torch.cuda.device(pc.device):
        ws = workspaces.get(lib.epc_knn_workspace_bytes(B, N))
        _lib.check(lib.epc_knn_dense(_ptr(pc), B, N, _lib.KNN_ARITH[arith], None, _ptr(dist), _ptr(ws), ws.numel(), _stream()))
    return dist


def knn(adj_matrix, k=20):
    """utils/tf_util.py:599-610: (B,N,M) pairwise distances -> (B,N,k) int32 indices of the k nearest
    (tf.nn.top_k(-adj) order: ascending distance, ties -> lower index)."""
    adj = as_cuda_f32(adj_matrix, "adj_matrix")
    lib = _lib.load()
    R = int(np.prod(adj.shape[:-1]))
    M = adj.shape[-1]
    idx = torch.empty(tuple(adj.shape[:-1]) + (k,), dtype=torch.int32, device=adj.device)
    with torch.cuda.device(adj.device):
        _lib.check(lib.epc_rows_topk_smallest(_ptr(adj), R, M, k, _ptr(idx), _stream()))
    return idx


def knn_graph(pc, arith="muladd", prune=True):
    """The fused path's kNN graph in original point order: (idx [B,N,20] int32 in tf.nn.top_k order,
    kth [B,N] fp32 = 20th largest a, count [B,N] int32 = size of the thresholded set)."""
    pc = _check_pc(pc)
    lib = _lib.load()
    B, N, _ = pc.shape
    idx = torch.empty((B, N, 20), dtype=torch.int32, device=pc.device)
    kth = torch.empty((B, N), dtype=torch.float32, device=pc.device)
    cnt = torch.empty((B, N), dtype=torch.int32, device=pc.device)
    fn = lib.epc_knn if prune else lib.epc_knn_noprune
    with torch.cuda.device(pc.device):
        ws = workspaces.get(lib.epc_knn_workspace_bytes(B, N))
        _lib.check(fn(_ptr(pc), B, N, _lib.KNN_ARITH[arith], _ptr(idx), _ptr(kth), _ptr(cnt), _ptr(ws), ws.numel(), _stream()))
    return idx, kth, cnt


def _dense(inputs2d, full_scope, cin, cout, bn, relu, store):
    store = store or variables.default_store()
    keep = []

    def arr(a):
        a = np.ascontiguousarray(a, dtype=np.float32)
        keep.append(a)
        return a.ctypes.data_as(_fp)

    if bn:
        m = "%s/bn/%s/bn/moments/Squeeze/ExponentialMovingAverage" % (full_scope, full_scope)
        v = "%s/bn/%s/bn/moments/Squeeze_1/ExponentialMovingAverage" % (full_scope, full_scope)
        bnp = _lib.EpcBN(arr(store[full_scope + "/bn/beta"]), arr(store[full_scope + "/bn/gamma"]), arr(store[m]), arr(store[v]))
    else:   # identity affine: gamma/sqrt(var+eps) == 1 exactly for var = 1-eps
        bnp = _lib.EpcBN(arr(np.zeros(cout)), arr(np.ones(cout)), arr(np.zeros(cout)), arr(np.full(cout, 1.0 - 1e-3)))
    w = store[full_scope + "/weights"]
    if w.size != cin * cout:
        raise ValueError("%s/weights has %d elements, expected %d x %d" % (full_scope, w.size, cin, cout))
    layer = _lib.EpcDense(arr(w), arr(store[full_scope + "/biases"]), bnp, cin, cout)
    x = as_cuda_f32(inputs2d, "inputs")
    y = torch.empty((x.shape[0], cout), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().epc_dense_forward(ctypes.byref(layer), _ptr(x), x.shape[0], _ptr(y), 1 if relu else 0, _stream()))
    return y


def conv1d(inputs, num_output_channels, kernel_size, scope, stride=1, padding="SAME", use_xavier=True, stddev=1e-3,
           weight_decay=0.0, activation_fn="relu", bn=False, bn_decay=None, is_training=None, variables_store=None):
    """utils/tf_util.py:52-107 for the only case on the path: kernel_size 1, stride 1.  inputs (B,L,C)."""
    _require_cuda()
    if kernel_size != 1 or stride != 1:
        raise NotImplementedError("only pointwise conv1d (kernel_size=1, stride=1) is on the EPC-Net path")
    if is_training:
        raise NotImplementedError("inference only (is_training=False)")
    B, L, C = inputs.shape
    full = (variables.current_scope() + "/" + scope).lstrip("/")
    y = _dense(inputs.reshape(B * L, C), full, C, num_output_channels, bn, activation_fn is not None, variables_store)
    return y.reshape(B, L, num_output_channels)


def fully_connected(inputs, num_outputs, scope, use_xavier=True, stddev=1e-3, weight_decay=0.0, activation_fn="relu",
                    bn=False, bn_decay=None, is_training=None, variables_store=None):
    """utils/tf_util.py:310-346.  inputs (B,C).  NOTE the default activation is ReLU, as in the reference."""
    _require_cuda()
    if is_training:
        raise NotImplementedError("inference only (is_training=False)")
    full = (variables.current_scope() + "/" + scope).lstrip("/")
    return _dense(inputs, full, inputs.shape[1], num_outputs, bn, activation_fn is not None, variables_store)


def max_pool2d(inputs, kernel_size, scope=None, stride=(2, 2), padding="VALID"):
    """utils/tf_util.py:349-372 for the case on the path: inputs (B,N,1,C), kernel [N,1] -> (B,1,1,C)
    (models/epc-net-l.py:91)."""
    _require_cuda()
    B, N, W, C = inputs.shape
    if W != 1 or list(kernel_size) != [N, 1] or padding != "VALID":
        raise NotImplementedError("only the global max-pool over points (kernel [num_points, 1], VALID) is implemented")
    x = as_cuda_f32(inputs, "inputs")
    y = torch.empty((B, 1, 1, C), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().epc_max_pool_points(_ptr(x), B, N, C, _ptr(y), _stream()))
    return y

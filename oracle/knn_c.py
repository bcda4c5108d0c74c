"""ctypes binding of oracle/knn_oracle.c (CPU ORACLE -- test infrastructure, not a product path)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libepc_oracle.so")
_lib = None
ARITH = {"muladd": 0, "fma": 1, 0: 0, 1: 1}


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "knn_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        fp = ctypes.POINTER(ctypes.c_float)
        ip = ctypes.POINTER(ctypes.c_int32)
        _lib.epc_oracle_knn.argtypes = [fp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ip, fp, ip]
        _lib.epc_oracle_knn.restype = ctypes.c_int
        _lib.epc_oracle_mask.argtypes = [fp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, fp, fp]
        _lib.epc_oracle_mask.restype = ctypes.c_int
    return _lib


def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _ip(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))


def knn(pc, k=20, arith="muladd"):
    """-> (idx [B,N,k] int32 in tf.nn.top_k order, kth [B,N] fp32 (value of a), count [B,N] int32)."""
    pc = np.ascontiguousarray(pc, dtype=np.float32)
    B, N, _ = pc.shape
    idx = np.empty((B, N, k), np.int32)
    kth = np.empty((B, N), np.float32)
    cnt = np.empty((B, N), np.int32)
    rc = lib().epc_oracle_knn(_fp(pc), B, N, k, ARITH[arith], _ip(idx), _fp(kth), _ip(cnt))
    if rc != 0:
        raise RuntimeError("epc_oracle_knn failed: %d" % rc)
    return idx, kth, cnt


def mask(pc, arith="muladd", want_a=False, k=20):
    """Dense (B,N,N) 0/1 fp32 mask of utils/tf_util.py:647-666 (and optionally ``a``)."""
    pc = np.ascontiguousarray(pc, dtype=np.float32)
    B, N, _ = pc.shape
    m = np.empty((B, N, N), np.float32)
    a = np.empty((B, N, N), np.float32) if want_a else None
    rc = lib().epc_oracle_mask(_fp(pc), B, N, k, ARITH[arith], _fp(a) if want_a else None, _fp(m))
    if rc != 0:
        raise RuntimeError("epc_oracle_mask failed: %d" % rc)
    return (m, a) if want_a else m

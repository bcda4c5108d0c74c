"""ctypes binding of libepc_b200.so (include/epc_b200.h).

There is no CPU fallback: importing this module without the built library raises, and every compute
entry point returns an error without a CUDA device.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_longlong, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libepc_b200.so")

EPC_OK, EPC_EINVAL, EPC_ECUDA, EPC_EWORKSPACE, EPC_EUNSUPPORTED = 0, -1, -2, -3, -4
ARCH_ENUM = {"epc-net": 0, "epc-net-l": 1, "kd_epc-net": 2, "kd_epc-net-l": 3}
KNN_ARITH = {"muladd": 0, "fma": 1}
POOLING = {"G_VLAD": 0, "NetVLAD": 1}


class EpcBN(Structure):
    _fields_ = [("beta_host", POINTER(c_float)), ("gamma_host", POINTER(c_float)),
                ("mean_host", POINTER(c_float)), ("var_host", POINTER(c_float))]


class EpcDense(Structure):
    _fields_ = [("weights_host", POINTER(c_float)), ("biases_host", POINTER(c_float)), ("bn", EpcBN),
                ("cin", c_int), ("cout", c_int)]


class EpcWeights(Structure):
    _fields_ = [("arch", c_int), ("knn_k", c_int), ("cluster_size", c_int), ("output_dim", c_int),
                ("groups", c_int), ("pooling", c_int), ("gating", c_int), ("n_blocks", c_int),
                ("conv", EpcDense * 12), ("conv5", EpcDense),
                ("cluster_weights_host", POINTER(c_float)), ("cluster_bn", EpcBN),
                ("cluster_weights2_host", POINTER(c_float)), ("hidden1_weights_host", POINTER(c_float)),
                ("hidden_bn", EpcBN), ("gating_weights_host", POINTER(c_float)), ("gating_bn", EpcBN),
                ("fc1", EpcDense)]


class EpcError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libepc_b200 error %d: %s" % (code, msg))
        self.code = code


# every symbol include/epc_b200.h declares: name -> (restype, argtypes)
_fp, _ip, _lp, _dp = POINTER(c_float), POINTER(c_int32), POINTER(c_int64), POINTER(c_double)
PROTOTYPES = {
    "epc_last_error": (c_char_p, []),
    "epc_abi_version": (c_int, []),
    "epc_launch_count": (c_longlong, []),
    "epc_launch_count_reset": (None, []),
    "epc_profile_enable": (None, [c_int]),
    "epc_profile_reset": (None, []),
    "epc_profile_read": (c_int, [c_int, POINTER(c_double), POINTER(c_longlong)]),
    "epc_stage_name": (c_char_p, [c_int]),
    "epc_microbench_ffma": (c_int, [c_int, POINTER(c_double), c_void_p]),
    "epc_set_device": (c_int, [c_int]),
    "epc_device_count": (c_int, []),
    "epc_knn_workspace_bytes": (c_size_t, [c_int, c_int]),
    "epc_knn": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "epc_knn_noprune": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "epc_knn_dense": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "epc_rows_topk_smallest": (c_int, [c_void_p, c_longlong, c_int, c_int, c_void_p, c_void_p]),
    "epc_dense_forward": (c_int, [POINTER(EpcDense), c_void_p, c_longlong, c_void_p, c_int, c_void_p]),
    "epc_max_pool_points": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "epc_model_create": (c_int, [POINTER(EpcWeights), POINTER(c_void_p)]),
    "epc_model_destroy": (None, [c_void_p]),
    "epc_embed_workspace_bytes": (c_size_t, [c_void_p, c_int, c_int]),
    "epc_embed": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "epc_vlad_workspace_bytes": (c_size_t, [c_void_p, c_int, c_int]),
    "epc_vlad_forward": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "epc_retrieve_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "epc_retrieve_topk": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_longlong, c_void_p, c_void_p,
                                  c_void_p, c_size_t, c_void_p]),
    "epc_retrieve_index_bytes": (c_size_t, [c_int, c_int]),
    "epc_retrieve_index_build": (c_int, [c_void_p, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "epc_retrieve_topk_indexed": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_longlong, c_void_p,
                                          c_void_p, c_void_p, c_size_t, c_void_p]),
    "epc_merge_topk": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "epc_merge_topk_strided": (c_int, [c_void_p, c_void_p, c_longlong, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "epc_radius_count": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_double, c_void_p, c_void_p]),
    "epc_radius_fill": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_double, c_void_p, c_void_p, c_void_p]),
}

_lib = None


def load():
    """Load libepc_b200.so (built by ``__graft_entry__.build()`` / ``make -C epc-net_b200/csrc``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s is missing: build it with `make -C %s` (or __graft_entry__.build()). "
                          "There is no CPU fallback." % (LIB_PATH, os.path.join(_HERE, "csrc")))
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError here == ABI mismatch with include/epc_b200.h
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise EpcError(rc, load().epc_last_error().decode("utf-8", "replace"))


STAGE_COUNT = 20


def profile_enable(on: bool):
    load().epc_profile_enable(1 if on else 0)


def profile_reset():
    load().epc_profile_reset()


def profile_read():
    """-> {stage name: (milliseconds, bracketed launches)} accumulated since the last reset."""
    lib = load()
    out = {}
    for s in range(STAGE_COUNT):
        ms, n = c_double(0), c_longlong(0)
        check(lib.epc_profile_read(s, ctypes.byref(ms), ctypes.byref(n)))
        if n.value:
            out[lib.epc_stage_name(s).decode()] = (ms.value, n.value)
    return out


def launch_count() -> int:
    return int(load().epc_launch_count())


def launch_count_reset():
    load().epc_launch_count_reset()

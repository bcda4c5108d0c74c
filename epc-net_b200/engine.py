"""Host-side engine: owns the device model handle (BN-folded weights), workspaces and chunking.

PyTorch is used for device memory, streams and pinned staging buffers only; all compute goes through
the C ABI of libepc_b200.so (``_lib``).
"""
from __future__ import annotations

import ctypes
import os
import threading

import numpy as np
import torch

from . import _lib, variables

_fp = ctypes.POINTER(ctypes.c_float)


def _require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("epc-net_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def as_cuda_f32(t, what="input", dtype=None):
    """A contiguous CUDA tensor of ``dtype`` (default fp32) for the C ABI, which takes raw device pointers and cannot check
    them: numpy arrays and host tensors are copied to the current device, other dtypes are converted (a float64 or half
    tensor passed by pointer would be silently reinterpreted)."""
    _require_cuda()
    dtype = torch.float32 if dtype is None else dtype
    if isinstance(t, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(t))
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch tensor or a numpy array, got %s" % (what, type(t).__name__))
    if not t.is_cuda:
        t = t.cuda()
    if t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


_POISON = int(os.environ["EPC_POISON_WORKSPACE"], 0) if os.environ.get("EPC_POISON_WORKSPACE") else None


class _Workspaces(object):
    """One grow-only byte buffer per (device, stream)."""

    def __init__(self):
        self._buf = {}
        self._lock = threading.Lock()

    def get(self, nbytes: int) -> torch.Tensor:
        key = (torch.cuda.current_device(), torch.cuda.current_stream().cuda_stream)
        with self._lock:
            b = self._buf.get(key)
            if b is None or b.numel() < nbytes:
                b = None
                self._buf.pop(key, None)
                b = torch.empty(int(nbytes * 1.05) + 4096, dtype=torch.uint8, device="cuda")
                self._buf[key] = b
            if _POISON is not None:      # debugging aid: EPC_POISON_WORKSPACE=<byte> refills the scratch before every call
                b.fill_(_POISON)
            return b

    def clear(self):
        with self._lock:
            self._buf.clear()


workspaces = _Workspaces()

_side_pool = {}
_side_lock = threading.Lock()


def side_streams(n: int):
    """The process-wide side streams of the current device (shared by every engine, so that the per-stream workspaces stay
    bounded however many engines come and go)."""
    dev = torch.cuda.current_device()
    with _side_lock:
        pool = _side_pool.setdefault(dev, [])
        while len(pool) < n:
            pool.append(torch.cuda.Stream())
        return pool[:n]


def _np_ptr(a):
    return a.ctypes.data_as(_fp)


class Engine(object):
    """A model instance on one device: ``arch`` + the variables found under ``scope`` in ``store``."""

    def __init__(self, arch: str, store: variables.VariableStore, scope: str, params: dict, device=None,
                 pooling: str = "G_VLAD", gating: bool = True, head_only: bool = False, vlad_prefix: str = None):
        _require_cuda()
        self.lib = _lib.load()
        self.arch = arch
        self.scope = scope
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        bscope, nblk, c5in, head = variables.arch_info(arch)
        self.n_blocks = nblk
        self.cluster_size = int(params.get("CLUSTER_SIZE", 64))
        self.output_dim = int(params.get("FEATURE_OUTPUT_DIM", 256))
        self.groups = int(params.get("GROUPS", 4))
        self.knn_k = int(params.get("KNN", 20))
        self.num_points = int(params.get("NUM_POINTS", 4096))
        self.is_vlad = head == "gvlad"
        # clouds per epc_embed call; unset: as large as 128, but at least `EMBED_STREAMS` chunks per call so that they overlap
        self.chunk = int(params["EMBED_CHUNK"]) if params.get("EMBED_CHUNK") else None
        # chunks of one call may alternate over this many side streams (own workspace each): the ALU-bound kNN of one chunk
        # then overlaps the HBM-bound head GEMMs of another
        self.nstreams = max(1, int(params.get("EMBED_STREAMS", os.environ.get("EPC_EMBED_STREAMS", 2))))
        self.knn_arith = _lib.KNN_ARITH[str(params.get("KNN_ARITH", "muladd"))]
        if int(params.get("INPUT_DIM", 3)) != 3:
            raise ValueError("INPUT_DIM must be 3 (xyz): conv1 of the shipped checkpoints is [1,3,64]")

        keep = []     # numpy arrays must outlive epc_model_create

        def arr(name):
            a = np.ascontiguousarray(store[name], dtype=np.float32)
            keep.append(a)
            return a

        def const(n, v):
            a = np.full((n,), v, np.float32)
            keep.append(a)
            return a

        def dummy_dense(cin, cout):     # head_only engines (stand-alone loupe API) never run the backbone
            bn = _lib.EpcBN(_np_ptr(const(cout, 0)), _np_ptr(const(cout, 1)), _np_ptr(const(cout, 0)), _np_ptr(const(cout, 1)))
            return _lib.EpcDense(_np_ptr(const(cin * cout, 0)), _np_ptr(const(cout, 0)), bn, cin, cout)

        def bn_template(full):      # tf_util.batch_norm_template naming (utils/tf_util.py:475-489)
            m = "%s/bn/%s/bn/moments/Squeeze/ExponentialMovingAverage" % (full, full)
            v = "%s/bn/%s/bn/moments/Squeeze_1/ExponentialMovingAverage" % (full, full)
            return _lib.EpcBN(_np_ptr(arr(full + "/bn/beta")), _np_ptr(arr(full + "/bn/gamma")), _np_ptr(arr(m)),
                              _np_ptr(arr(v)))

        def bn_slim(full):          # slim/contrib batch_norm naming (loupe.py:84,258,321)
            return _lib.EpcBN(_np_ptr(arr(full + "/beta")), _np_ptr(arr(full + "/gamma")),
                              _np_ptr(arr(full + "/moving_mean")), _np_ptr(arr(full + "/moving_variance")))

        def dense(full, cin, cout):
            w = arr(full + "/weights")
            if w.size != cin * cout:
                raise ValueError("%s/weights has %d elements, expected %d x %d" % (full, w.size, cin, cout))
            return _lib.EpcDense(_np_ptr(w), _np_ptr(arr(full + "/biases")), bn_template(full), cin, cout)

        w = _lib.EpcWeights()
        w.arch = _lib.ARCH_ENUM[arch]
        w.knn_k = self.knn_k
        w.cluster_size = self.cluster_size
        w.output_dim = self.output_dim
        w.groups = self.groups
        w.pooling = _lib.POOLING[pooling]
        w.gating = 1 if gating else 0
        w.n_blocks = nblk
        base = "%s/%s/" % (scope, bscope)
        cin = 3
        for i, cname in enumerate(variables.conv_names(arch)[:-1]):
            w.conv[i] = dummy_dense(cin, 64) if head_only else dense(base + cname, cin, 64)
            cin = 64
        w.conv5 = dummy_dense(c5in, 1024) if head_only else dense(base + "conv5", c5in, 1024)
        vs = (scope + "/VLAD/") if vlad_prefix is None else vlad_prefix
        if self.is_vlad:
            w.cluster_weights_host = _np_ptr(arr(vs + "cluster_weights"))
            w.cluster_bn = bn_slim(vs + "cluster_bn")
            w.cluster_weights2_host = _np_ptr(arr(vs + "cluster_weights2"))
            w.hidden1_weights_host = _np_ptr(arr(vs + "hidden1_weights"))
            w.hidden_bn = bn_slim(vs + "bn")
            if gating:
                w.gating_weights_host = _np_ptr(arr(vs + "gating_weights"))
                w.gating_bn = bn_slim(vs + "gating_bn")
        else:
            w.fc1 = dense(vs + "fc1", 1024, self.output_dim)
        handle = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.epc_set_device(self.device.index))
            _lib.check(self.lib.epc_model_create(ctypes.byref(w), ctypes.byref(handle)))
        self._h = handle
        del keep

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                self.lib.epc_model_destroy(h)
            except Exception:
                pass
            self._h = None

    def _chunk_for(self, n: int) -> int:
        if self.chunk:
            return self.chunk
        return max(1, min(256, -(-n // self.nstreams)))

    # ---- device-resident API ------------------------------------------------------------------
    def embed(self, xyz: torch.Tensor, want_feat: bool = False, out: torch.Tensor = None, single_call: bool = False):
        """xyz [B,N,3] fp32 CUDA tensor -> descriptors [B,D] (and KD features [B*N,1024]).  ``single_call``: one epc_embed on
        the current stream, no chunking and no side streams (embed_host pipelines its own chunks over its own streams)."""
        if not (xyz.is_cuda and xyz.dtype == torch.float32 and xyz.dim() == 3 and xyz.shape[-1] == 3):
            raise ValueError("embed expects a CUDA fp32 tensor [B,N,3], got %s %s %s" % (xyz.device, xyz.dtype, tuple(xyz.shape)))
        xyz = xyz.contiguous()
        B, N, _ = xyz.shape
        if out is None:
            out = torch.empty((B, self.output_dim), dtype=torch.float32, device=xyz.device)
        elif not (out.is_cuda and out.is_contiguous() and out.dtype == torch.float32 and tuple(out.shape) == (B, self.output_dim)):
            raise ValueError("out must be a contiguous CUDA fp32 tensor [B,%d]" % self.output_dim)
        feat = torch.empty((B * N, 1024), dtype=torch.float32, device=xyz.device) if want_feat else None
        with torch.cuda.device(xyz.device):
            chunk = max(1, B) if single_call else self._chunk_for(B)
            starts = list(range(0, B, chunk))
            fan = min(self.nstreams, len(starts))
            main = torch.cuda.current_stream()
            side = side_streams(fan) if fan > 1 else None
            if fan > 1:
                for st in side:
                    st.wait_stream(main)
            for ci, s in enumerate(starts):
                e = min(B, s + chunk)
                nb = e - s
                with torch.cuda.stream(side[ci % fan] if fan > 1 else main):
                    need = self.lib.epc_embed_workspace_bytes(self._h, nb, N)
                    ws = workspaces.get(need)              # one grow-only buffer per stream
                    f = feat[s * N:e * N] if want_feat else None
                    _lib.check(self.lib.epc_embed(self._h, _ptr(xyz[s:e]), nb, N, self.knn_arith, _ptr(out[s:e]), _ptr(f),
                                                  _ptr(ws), ws.numel(), _stream()))
            if fan > 1:
                for st in side:
                    main.wait_stream(st)
        return (out, feat) if want_feat else out

    def vlad(self, X: torch.Tensor, max_samples: int):
        """loupe forward on features X [B*max_samples, 1024] -> [B, D] (not L2-normalised)."""
        if not self.is_vlad:
            raise ValueError("%s has no VLAD head" % self.arch)
        X = as_cuda_f32(X, "reshaped_input")
        if X.dim() != 2 or X.shape[0] % max_samples or X.shape[1] != 1024:
            raise ValueError("reshaped_input must be [B*max_samples, 1024]")
        B = X.shape[0] // max_samples
        out = torch.empty((B, self.output_dim), dtype=torch.float32, device=X.device)
        with torch.cuda.device(X.device):
            chunk = self._chunk_for(B)
            for s in range(0, B, chunk):
                e = min(B, s + chunk)
                need = self.lib.epc_vlad_workspace_bytes(self._h, e - s, max_samples)
                ws = workspaces.get(need)
                _lib.check(self.lib.epc_vlad_forward(self._h, _ptr(X[s * max_samples:e * max_samples]), e - s, max_samples,
                                                     _ptr(out[s:e]), _ptr(ws), ws.numel(), _stream()))
        return out

    # ---- host-buffer API (the reference's feed_dict / fetch path, evaluate.py:378-390) ---------------
    def embed_host(self, clouds: np.ndarray, out: np.ndarray = None, pipeline_chunk: int = None) -> np.ndarray:
        """clouds [n,N,3] host array -> [n,D] host array.  H2D of chunk i+1 overlaps compute of chunk i."""
        clouds = np.ascontiguousarray(clouds, dtype=np.float32)
        n, N, _ = clouds.shape
        chunk = pipeline_chunk or self._chunk_for(n)
        D = self.output_dim
        if out is None:
            out = np.empty((n, D), np.float32)
        if n == 0:
            return out
        with torch.cuda.device(self.device):
            starts = list(range(0, n, chunk))
            fan = min(self.nstreams, len(starts))          # compute streams the chunks alternate over
            slots = fan + 1                                # staging slots: one being filled while `fan` are in use
            # staging buffers are cached: pinning host memory costs far more than the copy itself
            st = getattr(self, "_staging", None)
            if st is None or st["chunk"] < chunk or st["N"] != N or st["n"] < n or st["slots"] < slots:
                st = {"chunk": chunk, "N": N, "n": n, "slots": slots,
                      "pin_in": [torch.empty((chunk, N, 3), dtype=torch.float32).pin_memory() for _ in range(slots)],
                      "pin_out": torch.empty((n, D), dtype=torch.float32).pin_memory(),
                      "dev_in": [torch.empty((chunk, N, 3), dtype=torch.float32, device=self.device) for _ in range(slots)],
                      "dev_out": torch.empty((n, D), dtype=torch.float32, device=self.device),
                      "copy_stream": torch.cuda.Stream()}
                self._staging = st
            pin_in, dev_in, copy_stream = st["pin_in"], st["dev_in"], st["copy_stream"]
            pin_out, dev_out = st["pin_out"][:n], st["dev_out"][:n]
            main = torch.cuda.current_stream()
            compute = side_streams(fan) if fan > 1 else [main]
            copy_stream.wait_stream(main)
            for cs in compute:
                if cs is not main:
                    cs.wait_stream(main)
            in_ready = [torch.cuda.Event() for _ in range(slots)]
            in_free = [torch.cuda.Event() for _ in range(slots)]
            host_free = [None] * slots
            for i, s in enumerate(starts):
                e = min(n, s + chunk)
                slot = i % slots
                cs = compute[i % fan]
                if host_free[slot] is not None:
                    host_free[slot].synchronize()        # the H2D that last read this pinned slot is done
                pin_in[slot][:e - s].copy_(torch.from_numpy(clouds[s:e]))
                with torch.cuda.stream(copy_stream):
                    if i >= slots:
                        copy_stream.wait_event(in_free[slot])      # compute of chunk i-slots released the device slot
                    dev_in[slot][:e - s].copy_(pin_in[slot][:e - s], non_blocking=True)
                    in_ready[slot].record(copy_stream)
                    host_free[slot] = in_ready[slot]
                cs.wait_event(in_ready[slot])
                with torch.cuda.stream(cs):
                    self.embed(dev_in[slot][:e - s], out=dev_out[s:e], single_call=True)       # one library call on `cs`
                in_free[slot].record(cs)
            for cs in compute:
                if cs is not main:
                    main.wait_stream(cs)
            pin_out.copy_(dev_out, non_blocking=True)
            main.synchronize()
        out[...] = pin_out.numpy()
        return out


# ---- engine cache keyed like a TF graph: (store, scope, arch, head options, device) -------------------
_engines = {}
_engines_lock = threading.Lock()
_ENGINE_CACHE_MAX = 16        # each engine owns a device weight blob (~6 MB)


def get_engine(arch: str, params: dict, scope: str = None, store: variables.VariableStore = None,
               pooling: str = "G_VLAD", gating: bool = True) -> Engine:
    _require_cuda()
    scope = variables.current_scope("query_triplets") if scope is None else scope
    store = variables.default_store() if store is None else store
    key = (store.uid, store.version, scope, arch, pooling, gating, torch.cuda.current_device(),
           int(params.get("CLUSTER_SIZE", 64)), int(params.get("FEATURE_OUTPUT_DIM", 256)), int(params.get("GROUPS", 4)),
           int(params.get("KNN", 20)), str(params.get("KNN_ARITH", "muladd")), int(params.get("EMBED_CHUNK") or 0),
           int(params.get("EMBED_STREAMS", os.environ.get("EPC_EMBED_STREAMS", 2))))
    with _engines_lock:
        eng = _engines.pop(key, None)
        if eng is None:
            eng = Engine(arch, store, scope, params, pooling=pooling, gating=gating)
            while len(_engines) >= _ENGINE_CACHE_MAX:         # least recently used first (dicts keep insertion order)
                _engines.pop(next(iter(_engines)))
        _engines[key] = eng                                   # (re-)insert as the most recently used
        return eng


def clear_engines():
    with _engines_lock:
        _engines.clear()
    workspaces.clear()

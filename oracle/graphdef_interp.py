"""CPU ORACLE tooling (test infrastructure, NOT a product path): a small numpy interpreter for the
TensorFlow-1.12 GraphDefs the reference ships in ``exp/*/saved_model/*.ckpt.meta``.

Why: TensorFlow cannot be installed here, but the reference's serialised graphs can
be parsed with the TF protobufs bundled in ``tensorboard``.  Executing the *reference's own graph*
op by op -- instead of trusting a hand restatement of the Python source -- is the strongest pin
available for ``oracle/epc_oracle.py``.  ``tests/golden/make_golden.py`` uses this module (in the
build container only: it reads /root/reference) to mint the committed ``tests/golden/graph_*.npz``.

Semantics implemented: exactly the op set on the inference path (lazy evaluation from the fetch;
``Switch``/``Merge`` dead-branch propagation implements ``tf.cond(is_training, ...)``).  Variables
are looked up by node name in a ``name -> ndarray`` dict (the checkpoint names).

Arithmetic caveat: every op is IEEE fp32, but the summation order *inside* MatMul/Conv2D is the
host BLAS's, not Eigen's/cuBLAS's.  The only place where that matters for set-valued results is the
K=3 ``BatchMatMul`` feeding TopKV2; ``small_k_matmul`` selects the reconstruction
used there ("muladd" = separately rounded ((x x'+y y')+z z'), the TF-1.12 CPU wheel's AVX/no-FMA
Eigen order; "blas" = whatever numpy's BLAS does).
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


class _Dead(object):
    def __repr__(self):
        return "<DEAD>"


DEAD = _Dead()

_DT = {1: np.float32, 2: np.float64, 3: np.int32, 9: np.int64, 10: np.bool_}


def load_meta_graph(path):
    from tensorboard.compat.proto import meta_graph_pb2
    mg = meta_graph_pb2.MetaGraphDef()
    with open(path, "rb") as f:
        mg.ParseFromString(f.read())
    return mg


def _tensor_proto_to_numpy(tp):
    dtype = _DT[tp.dtype]
    shape = [d.size for d in tp.tensor_shape.dim]
    if tp.tensor_content:
        return np.frombuffer(tp.tensor_content, dtype=dtype).reshape(shape).copy()
    if dtype == np.float32:
        vals = list(tp.float_val)
    elif dtype == np.float64:
        vals = list(tp.double_val)
    elif dtype == np.int32:
        vals = list(tp.int_val)
    elif dtype == np.int64:
        vals = list(tp.int64_val)
    elif dtype == np.bool_:
        vals = list(tp.bool_val)
    else:  # pragma: no cover
        raise NotImplementedError(tp.dtype)
    n = int(np.prod(shape)) if shape else 1
    if len(vals) == 0:
        arr = np.zeros(n, dtype)
    elif len(vals) == 1:
        arr = np.full(n, vals[0], dtype)
    else:
        arr = np.asarray(vals, dtype)
    return arr.reshape(shape)


class GraphInterpreter(object):
    def __init__(self, graph_def, variables, small_k_matmul="muladd"):
        self.nodes = {n.name: n for n in graph_def.node}
        self.variables = variables
        self.small_k_matmul = small_k_matmul
        self.cache = {}
        self.executed_ops = []

    # ---- public -------------------------------------------------------------------------------
    def run(self, fetches, feed_dict):
        self.cache = {}
        for k, v in feed_dict.items():
            self.cache[self._canon(k)] = v
        single = isinstance(fetches, str)
        outs = [self._eval(self._canon(f)) for f in ([fetches] if single else fetches)]
        return outs[0] if single else outs

    def free(self, *names):
        for n in names:
            self.cache.pop(self._canon(n), None)

    # ---- internals ----------------------------------------------------------------------------
    @staticmethod
    def _canon(name):
        if name.startswith("^"):
            raise ValueError("control input")
        return name if ":" in name else name + ":0"

    def _inputs(self, node):
        return [i for i in node.input if not i.startswith("^")]

    def _eval(self, tname):
        if tname in self.cache:
            return self.cache[tname]
        nname, oidx = tname.rsplit(":", 1)
        oidx = int(oidx)
        node = self.nodes[nname]
        outs = self._exec(node)
        if not isinstance(outs, tuple):
            outs = (outs,)
        for i, o in enumerate(outs):
            self.cache["%s:%d" % (nname, i)] = o
        return outs[oidx]

    def _in(self, node, i):
        return self._eval(self._canon(self._inputs(node)[i]))

    def _exec(self, node):
        op = node.op
        ins = self._inputs(node)
        self.executed_ops.append(op)
        if op == "Merge":
            for i in range(len(ins)):
                v = self._in(node, i)
                if v is not DEAD:
                    return (v, np.int32(i))
            return (DEAD, DEAD)
        if op == "Switch":
            pred = self._in(node, 1)
            if pred is DEAD:
                return (DEAD, DEAD)
            # evaluate data only on the live side (both outputs carry the same data)
            data = self._in(node, 0)
            return (DEAD, data) if bool(pred) else (data, DEAD)
        # generic ops: any dead input kills the op (evaluate left to right, stop at the first)
        vals = []
        for i in range(len(ins)):
            v = self._in(node, i)
            if v is DEAD:
                return tuple([DEAD] * 3)
            vals.append(v)
        fn = getattr(self, "_op_" + op, None)
        if fn is None:
            raise NotImplementedError("op %s (node %s) is not on the supported inference path" % (op, node.name))
        return fn(node, *vals)

    # ---- ops ----------------------------------------------------------------------------------
    def _op_Placeholder(self, node):
        raise KeyError("placeholder %s was not fed" % node.name)

    def _op_Const(self, node):
        return _tensor_proto_to_numpy(node.attr["value"].tensor)

    def _op_VariableV2(self, node):
        return np.asarray(self.variables[node.name])

    def _op_Identity(self, node, x):
        return x

    _op_StopGradient = _op_Identity

    def _op_Reshape(self, node, x, shape):
        return np.reshape(x, [int(s) for s in shape])

    def _op_ConcatV2(self, node, *args):
        return np.concatenate(args[:-1], axis=int(args[-1]))

    def _op_Transpose(self, node, x, perm):
        return np.transpose(x, [int(p) for p in perm])

    def _op_ExpandDims(self, node, x, dim):
        return np.expand_dims(x, int(dim))

    def _op_Squeeze(self, node, x):
        dims = list(node.attr["squeeze_dims"].list.i)
        return np.squeeze(x, axis=tuple(int(d) for d in dims)) if dims else np.squeeze(x)

    def _matmul(self, a, b):
        if a.shape[-1] <= 4 and self.small_k_matmul == "muladd":
            acc = a[..., :, 0:1] * b[..., 0:1, :]
            for kk in range(1, a.shape[-1]):
                acc = acc + a[..., :, kk:kk + 1] * b[..., kk:kk + 1, :]
            return acc.astype(F32)
        return np.matmul(a, b)

    def _op_BatchMatMul(self, node, a, b):
        if node.attr["adj_x"].b:
            a = np.swapaxes(a, -1, -2)
        if node.attr["adj_y"].b:
            b = np.swapaxes(b, -1, -2)
        return self._matmul(a, b)

    def _op_MatMul(self, node, a, b):
        if node.attr["transpose_a"].b:
            a = a.T
        if node.attr["transpose_b"].b:
            b = b.T
        return self._matmul(a, b)

    def _op_Conv2D(self, node, x, w):
        assert node.attr["data_format"].s in (b"NHWC", b""), node.attr["data_format"].s
        assert list(node.attr["strides"].list.i) == [1, 1, 1, 1]
        kh, kw, cin, cout = w.shape
        assert kh == 1 and kw == 1, "only 1x1 convolutions are on the path"
        return np.matmul(x, w.reshape(cin, cout))

    def _op_BiasAdd(self, node, x, b):
        return x + b

    def _op_Mul(self, node, a, b):
        return a * b

    def _op_Add(self, node, a, b):
        return a + b

    def _op_Sub(self, node, a, b):
        return a - b

    def _op_RealDiv(self, node, a, b):
        return a / b

    def _op_Neg(self, node, a):
        return -a

    def _op_Square(self, node, a):
        return a * a

    def _op_Rsqrt(self, node, a):
        return (F32(1.0) / np.sqrt(a)).astype(a.dtype)

    def _op_Maximum(self, node, a, b):
        return np.maximum(a, b)

    def _op_Relu(self, node, a):
        return np.maximum(a, F32(0))

    def _op_Sigmoid(self, node, a):
        return (F32(1.0) / (F32(1.0) + np.exp(-a))).astype(a.dtype)

    def _op_Softmax(self, node, a):
        m = np.max(a, axis=-1, keepdims=True)
        e = np.exp(a - m)
        return (e / np.sum(e, axis=-1, keepdims=True, dtype=a.dtype)).astype(a.dtype)

    def _reduce(self, node, fn, x, axes):
        axes = tuple(int(a) for a in np.atleast_1d(axes))
        keep = bool(node.attr["keep_dims"].b)
        return fn(x, axis=axes, keepdims=keep)

    def _op_Sum(self, node, x, axes):
        return self._reduce(node, lambda v, **kw: np.sum(v, dtype=v.dtype, **kw), x, axes)

    def _op_Min(self, node, x, axes):
        return self._reduce(node, np.min, x, axes)

    def _op_Max(self, node, x, axes):
        return self._reduce(node, np.max, x, axes)

    def _op_TopKV2(self, node, x, k):
        k = int(k)
        n = x.shape[-1]
        # values only are consumed on the path (Min over them); sorted=true order is irrelevant to Min
        part = np.partition(x, n - k, axis=-1)[..., n - k:]
        vals = -np.sort(-part, axis=-1)
        return (vals, DEAD)

    def _op_GreaterEqual(self, node, a, b):
        return a >= b

    def _op_Cast(self, node, a):
        return a.astype(_DT[node.attr["DstT"].type])

    def _op_FusedBatchNorm(self, node, x, scale, offset, mean, var):
        assert not node.attr["is_training"].b
        eps = F32(node.attr["epsilon"].f)
        inv = (F32(1.0) / np.sqrt(var + eps)).astype(F32)
        # NHWC, channel last
        return ((x - mean) * inv * scale + offset, mean, var, DEAD, DEAD)

    def _op_MaxPool(self, node, x):
        ks = list(node.attr["ksize"].list.i)
        st = list(node.attr["strides"].list.i)
        assert node.attr["padding"].s == b"VALID"
        assert ks[0] == 1 and ks[3] == 1 and ks[2] == 1 and ks[1] == x.shape[1], (ks, x.shape)
        return np.max(x, axis=1, keepdims=True)

    def _op_SplitV(self, node, x, sizes, axis):
        idx = np.cumsum([int(s) for s in sizes])[:-1]
        return tuple(np.split(x, idx, axis=int(axis)))

    def _op_Pack(self, node, *xs):
        return np.stack(xs, axis=int(node.attr["axis"].i))


def find_placeholders(graph_def):
    """[(name, shape or None)] in graph order."""
    out = []
    for n in graph_def.node:
        if n.op == "Placeholder":
            sh = n.attr["shape"].shape
            out.append((n.name, [d.size for d in sh.dim] if not sh.unknown_rank else None))
    return out

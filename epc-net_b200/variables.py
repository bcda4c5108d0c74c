"""Variable store for the EPC-Net embedding path: TensorFlow variable names -> fp32 arrays.

The reference keeps its weights as graph variables restored *by name* with ``tf.train.Saver``
(evaluate.py:253,263-269).  This module keeps the same names (taken from the shipped
``exp/*/saved_model/*.ckpt.index`` files) so that

* a real checkpoint (``.index`` + ``.data-00000-of-00001``) can be loaded without TensorFlow
  (``tf_bundle.read_checkpoint``), and
* seeded synthetic weights with the reference's initialiser distributions can stand in for the
  missing ``.data`` blobs (utils/tf_util.py:41-42, loupe.py:249-253,278-282,314-316,75-79).

Pure numpy; no device code here.
"""
from __future__ import annotations

import contextlib
import threading
import itertools
from collections import OrderedDict

import numpy as np

ARCHS = ("epc-net", "epc-net-l", "kd_epc-net", "kd_epc-net-l")

# (backbone scope, number of ProxyConv blocks, conv5 input channels, head kind)
_ARCH_INFO = {
    "epc-net": ("fastdgcnn", 4, 256, "gvlad"),        # models/epc-net.py:62,134-139,141-149
    "epc-net-l": ("fastdgcnn", 2, 128, "maxfc"),      # models/epc-net-l.py:44,82-95
    "kd_epc-net": ("fastdgcnn", 4, 256, "gvlad"),     # models/kd_epc-net.py (teacher)
    "kd_epc-net-l": ("BACKBONE", 2, 128, "maxfc"),    # models/kd_epc-net-l.py:44 (student)
}


def arch_info(arch: str):
    if arch not in _ARCH_INFO:
        raise ValueError("unknown ARCH %r (expected one of %s)" % (arch, ", ".join(ARCHS)))
    return _ARCH_INFO[arch]


def conv_names(arch: str):
    """Names of the 64-wide pointwise conv layers in execution order, then 'conv5'."""
    _, nblk, _, _ = arch_info(arch)
    names = []
    for b in range(1, nblk + 1):
        names += ["conv%d" % b, "conv%d_a" % b, "conv%d_b" % b]
    return names + ["conv5"]


def _ema_names(full_scope: str):
    # utils/tf_util.py:475-489: ema.average(batch_mean) lives under <scope>/bn/<full scope path>/bn/moments/...
    base = "%s/bn/%s/bn/moments/" % (full_scope, full_scope)
    return base + "Squeeze/ExponentialMovingAverage", base + "Squeeze_1/ExponentialMovingAverage"


def variable_specs(arch: str, scope: str = "query_triplets", cluster_size: int = 64,
                   output_dim: int = 256, groups: int = 4, pooling: str = "G_VLAD"):
    """OrderedDict name -> shape for every inference variable of ``arch`` under ``scope``."""
    bscope, nblk, c5in, head = arch_info(arch)
    specs = OrderedDict()
    cin = 3
    for cname in conv_names(arch):
        cout = 1024 if cname == "conv5" else 64
        if cname == "conv5":
            cin = c5in
        full = "%s/%s/%s" % (scope, bscope, cname)
        specs[full + "/weights"] = (1, cin, cout)
        specs[full + "/biases"] = (cout,)
        specs[full + "/bn/beta"] = (cout,)
        specs[full + "/bn/gamma"] = (cout,)
        m, v = _ema_names(full)
        specs[m] = (cout,)
        specs[v] = (cout,)
        cin = 64
    v = scope + "/VLAD/"
    if head == "gvlad":
        hid_in = 1024 * cluster_size // groups if pooling == "G_VLAD" else 1024 * cluster_size
        specs[v + "cluster_weights"] = (1024, cluster_size)
        for s in ("beta", "gamma", "moving_mean", "moving_variance"):
            specs[v + "cluster_bn/" + s] = (cluster_size,)
        specs[v + "cluster_weights2"] = (1, 1024, cluster_size)
        specs[v + "hidden1_weights"] = (hid_in, output_dim)
        for s in ("beta", "gamma", "moving_mean", "moving_variance"):
            specs[v + "bn/" + s] = (output_dim,)
        specs[v + "gating_weights"] = (output_dim, output_dim)
        for s in ("beta", "gamma", "moving_mean", "moving_variance"):
            specs[v + "gating_bn/" + s] = (output_dim,)
    else:
        full = v + "fc1"
        specs[full + "/weights"] = (1024, output_dim)
        specs[full + "/biases"] = (output_dim,)
        specs[full + "/bn/beta"] = (output_dim,)
        specs[full + "/bn/gamma"] = (output_dim,)
        m, vv = _ema_names(full)
        specs[m] = (output_dim,)
        specs[vv] = (output_dim,)
    return specs


def synthetic_variables(arch: str, seed: int = 0, scope: str = "query_triplets", **kw):
    """Seeded stand-in weights (SURVEY.md section 8d "Weights").

    Distributions follow the reference initialisers; the BN statistics are made non-trivial so the
    folded affine is exercised.  cluster_bn's moving variance is set near the actual variance of
    the pre-BN logits (~1/1024 for unit-norm inputs) so the soft assignment is not degenerate.
    """
    rng = np.random.default_rng(seed)
    out = OrderedDict()
    for name, shape in variable_specs(arch, scope, **kw).items():
        leaf = name.rsplit("/", 1)[-1]
        if leaf == "weights":
            fan_in, fan_out = shape[-2], shape[-1]
            lim = np.sqrt(6.0 / (fan_in + fan_out))           # xavier_initializer (uniform)
            a = rng.uniform(-lim, lim, shape)
        elif leaf == "biases":
            a = rng.normal(0.0, 0.05, shape)
        elif leaf in ("cluster_weights", "cluster_weights2"):
            a = rng.normal(0.0, 1.0 / np.sqrt(1024.0), shape)
        elif leaf == "hidden1_weights":
            a = rng.normal(0.0, 1.0 / np.sqrt(kw.get("cluster_size", 64)), shape)
        elif leaf == "gating_weights":
            a = rng.normal(0.0, 1.0 / np.sqrt(shape[0]), shape)
        elif leaf == "gamma":
            a = rng.uniform(0.5, 1.5, shape)
        elif leaf in ("beta", "moving_mean") or name.endswith("Squeeze/ExponentialMovingAverage"):
            a = rng.normal(0.0, 0.1, shape)
            if "cluster_bn" in name:
                a = a * 0.03
        elif leaf == "moving_variance" or name.endswith("Squeeze_1/ExponentialMovingAverage"):
            a = rng.uniform(0.5, 1.5, shape)
            if "cluster_bn" in name:
                a = a / 1024.0
        else:  # pragma: no cover
            raise AssertionError(name)
        out[name] = np.ascontiguousarray(a, dtype=np.float32)
    return out


class VariableStore(object):
    """name -> np.float32 array; the stand-in for the TF variable collection + Saver."""

    _uids = itertools.count(1)

    def __init__(self, values=None):
        self._v = OrderedDict()
        self.uid = next(VariableStore._uids)     # process-unique, never reused (id() is, once a store is collected)
        self.version = 0
        if values:
            self.update(values)

    def update(self, values):
        for k, a in values.items():
            self._v[k] = np.ascontiguousarray(a, dtype=np.float32)
        self.version += 1

    def restore(self, ckpt_prefix: str):
        """Equivalent of ``saver.restore(sess, path)`` (evaluate.py:263-269), without TensorFlow."""
        from . import tf_bundle
        vals = tf_bundle.read_checkpoint(ckpt_prefix, skip_optimizer_slots=True)
        self.update({k: v for k, v in vals.items() if v.dtype == np.float32})

    def __contains__(self, k):
        return k in self._v

    def __getitem__(self, k):
        try:
            return self._v[k]
        except KeyError:
            raise KeyError("variable %r not in store (have %d variables; was a checkpoint restored "
                           "or init_synthetic() called for this scope/ARCH?)" % (k, len(self._v)))

    def keys(self):
        return self._v.keys()

    def items(self):
        return self._v.items()


_default_store = VariableStore()
_scope_local = threading.local()


def default_store() -> VariableStore:
    return _default_store


def init_synthetic(arch: str, seed: int = 0, scope: str = None, store: VariableStore = None, **kw):
    """Fill ``store`` (default: the global store) with seeded weights for ``arch``."""
    scope = current_scope("query_triplets") if scope is None else scope
    store = _default_store if store is None else store
    store.update(synthetic_variables(arch, seed, scope, **kw))
    return store


@contextlib.contextmanager
def variable_scope(name: str):
    """Mirror of ``with tf.variable_scope(name)`` as used at evaluate.py:248, kd_evaluate.py:256."""
    stack = getattr(_scope_local, "stack", None)
    if stack is None:
        stack = _scope_local.stack = []
    stack.append(name)
    try:
        yield "/".join(stack)
    finally:
        stack.pop()


def current_scope(default: str = "") -> str:
    stack = getattr(_scope_local, "stack", None)
    return "/".join(stack) if stack else default

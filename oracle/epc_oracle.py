"""CPU ORACLE (test infrastructure, NOT a product path) -- numpy restatement of the reference's
embedding path, dense-as-written.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module; the product (``epc-net_b200/``) never does.

Parity status: the reference's TensorFlow 1.12 runtime is not runnable here and its checkpoint
``.data`` blobs are absent, so this restatement is pinned against
(a) the reference's own *serialised graph* (``exp/*/saved_model/*.ckpt.meta``) executed op by op by
    ``oracle/graphdef_interp.py`` -- goldens in ``tests/golden/graph_*.npz`` -- and
(b) ``sklearn.neighbors.KDTree`` (the library evaluate.py:463,481 calls) for retrieval.
What remains UNPINNED: TensorFlow's internal fp32 summation order inside MatMul/Conv2D kernels
(affects the last bits, and for the kNN mask the set of 1 row in ~4096).  The kNN
arithmetic is therefore a documented reconstruction: see ``oracle/knn_oracle.c``.

Every function cites the reference lines it follows (paths relative to /root/reference).
All arithmetic is fp32 unless noted.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32
BN_EPS = F32(1e-3)     # tf.nn.batch_normalization(..., 1e-3) utils/tf_util.py:490; slim/contrib default 0.001
L2_EPS = F32(1e-12)    # tf.nn.l2_normalize default epsilon


# ------------------------------------------------------------------------------------------------
# utils/tf_util.py
# ------------------------------------------------------------------------------------------------
def pairwise_a(pc, arith="muladd"):
    """``a = -(|p_i|^2 + (-2 p_i.p_j) + |p_j|^2)``  -- utils/tf_util.py:651-656.

    ``arith='muladd'``: inner product as separately rounded ``((x x' + y y') + z z')`` (the stock
    TF-1.12 CPU wheel: AVX, no FMA).  ``arith='fma'`` (cuBLAS-style ``fma(z,z',fma(y,y',x*x'))``) is
    only implemented bit-exactly in the C oracle (numpy has no fused multiply-add).
    """
    if arith != "muladd":
        raise NotImplementedError("numpy oracle implements arith='muladd' only; use oracle.knn_c")
    pc = np.asarray(pc, dtype=F32)
    x, y, z = pc[..., 0], pc[..., 1], pc[..., 2]
    inner = (x[:, :, None] * x[:, None, :] + y[:, :, None] * y[:, None, :]) + z[:, :, None] * z[:, None, :]
    inner2 = F32(-2.0) * inner                                  # :653
    sq = (x * x + y * y) + z * z                                # :654 reduce_sum(square(pc), -1)
    a = -((sq[:, :, None] + inner2) + sq[:, None, :])           # :656 left-to-right
    return a


def pairwise_distance(pc, arith="muladd"):
    """utils/tf_util.py:577-596 (same expression without the final negation)."""
    return -pairwise_a(pc, arith)


def topk_threshold(a, k=20):
    """``kth = reduce_min(top_k(a, 20).values, 2)`` -- utils/tf_util.py:660-663."""
    # k-th largest along the last axis == element (n-k) of an ascending partition
    n = a.shape[-1]
    part = np.partition(a, n - k, axis=-1)
    return part[..., n - k][..., None]


def pairwise_distance_mask(pc, k=20, arith="muladd"):
    """utils/tf_util.py:647-666.  NOTE the literal 20 at :660: the ``k`` argument is ignored."""
    a = pairwise_a(pc, arith)
    kth = topk_threshold(a, 20)
    return (a >= kth).astype(F32)                               # :664-665


def knn(adj_matrix, k=20):
    """utils/tf_util.py:599-610: indices of the k smallest entries per row (tf.nn.top_k order:
    descending value of -adj, ties -> lower index first)."""
    neg = -np.asarray(adj_matrix, dtype=F32)
    order = np.argsort(-neg, axis=-1, kind="stable")            # stable => lower index wins ties
    return order[..., :k].astype(np.int32)


def batch_norm_inference(x, beta, gamma, mean, var):
    """tf.nn.batch_normalization with EMA statistics -- utils/tf_util.py:486-490; node order in the
    shipped GraphDef: batchnorm/add, Rsqrt, mul, mul_1, mul_2, sub, add_1."""
    inv = (F32(1.0) / np.sqrt(var + BN_EPS)).astype(F32) * gamma
    return x * inv + (beta - mean * inv)


def fused_batch_norm_inference(x, beta, gamma, mean, var):
    """FusedBatchNorm(is_training=False), used by tf.contrib.layers.batch_norm / slim.batch_norm on
    rank-2 inputs with the default fused=None (loupe.py:84-89, 321): (x-mean)*rsqrt(var+eps)*gamma+beta."""
    inv = (F32(1.0) / np.sqrt(var + BN_EPS)).astype(F32)
    return (x - mean) * inv * gamma + beta


def _conv_vars(V, full):
    m = "%s/bn/%s/bn/moments/Squeeze/ExponentialMovingAverage" % (full, full)
    v = "%s/bn/%s/bn/moments/Squeeze_1/ExponentialMovingAverage" % (full, full)
    return (V[full + "/weights"], V[full + "/biases"], V[full + "/bn/beta"], V[full + "/bn/gamma"], V[m], V[v])


def conv1d(x, V, full):
    """tf_util.conv1d(..., kernel 1, bn=True, activation relu) -- utils/tf_util.py:85-107."""
    W, b, beta, gamma, mean, var = _conv_vars(V, full)
    y = np.matmul(x, W[0]) + b                                  # tf.nn.conv1d + bias_add :94-99
    y = batch_norm_inference(y, beta, gamma, mean, var)         # :101-103
    return np.maximum(y, F32(0.0))                              # :105-106


def fully_connected(x, V, full):
    """tf_util.fully_connected(bn=True) with its DEFAULT relu -- utils/tf_util.py:310-346."""
    W, b, beta, gamma, mean, var = _conv_vars(V, full)
    y = np.matmul(x, W) + b
    y = batch_norm_inference(y, beta, gamma, mean, var)
    return np.maximum(y, F32(0.0))


def l2_normalize(x, axis):
    """tf.nn.l2_normalize: x * rsqrt(max(sum(x^2), 1e-12))."""
    ss = np.sum(x * x, axis=axis, keepdims=True, dtype=F32)
    return x * (F32(1.0) / np.sqrt(np.maximum(ss, L2_EPS))).astype(F32)


# ------------------------------------------------------------------------------------------------
# loupe.py
# ------------------------------------------------------------------------------------------------
def _softmax(x):
    m = np.max(x, axis=-1, keepdims=True)
    e = np.exp(x - m)
    return (e / np.sum(e, axis=-1, keepdims=True, dtype=F32)).astype(F32)


def context_gating(v, V, vs):
    """PoolingBaseModel.context_gating -- loupe.py:61-101 (add_batch_norm=True branch)."""
    gates = np.matmul(v, V[vs + "gating_weights"])                              # :80
    gates = fused_batch_norm_inference(gates, V[vs + "gating_bn/beta"], V[vs + "gating_bn/gamma"],
                                       V[vs + "gating_bn/moving_mean"], V[vs + "gating_bn/moving_variance"])
    gates = (F32(1.0) / (F32(1.0) + np.exp(-gates))).astype(F32)                # :97
    return v * gates                                                            # :99


def vlad_forward(X, V, vs, max_samples, cluster_size=64, output_dim=256, groups=4, gating=True,
                 pooling="G_VLAD", return_intermediate=False):
    """G_VLAD.forward (loupe.py:233-333) / NetVLAD.forward (loupe.py:119-214).

    X: (B*max_samples, feature_size) -- the caller has already L2-normalised rows (models/epc-net.py:147-148).
    """
    X = np.asarray(X, dtype=F32)
    F = X.shape[1]
    act = np.matmul(X, V[vs + "cluster_weights"])                               # :255 / :141
    act = batch_norm_inference(act, V[vs + "cluster_bn/beta"], V[vs + "cluster_bn/gamma"],
                               V[vs + "cluster_bn/moving_mean"], V[vs + "cluster_bn/moving_variance"])  # fused=False
    act = _softmax(act)                                                         # :272
    act = act.reshape(-1, max_samples, cluster_size)                            # :274
    a_sum = np.sum(act, axis=-2, keepdims=True, dtype=F32)                      # :276
    a = a_sum * V[vs + "cluster_weights2"]                                      # :284  (B,F,K)
    actT = np.transpose(act, (0, 2, 1))                                         # :286
    Xr = X.reshape(-1, max_samples, F)                                          # :288
    vlad = np.matmul(actT, Xr)                                                  # :290  (B,K,F)
    vlad = np.transpose(vlad, (0, 2, 1))                                        # :291  (B,F,K)
    vlad = vlad - a                                                             # :292
    vlad = l2_normalize(vlad, 1)                                                # :295 intra-norm over F
    vlad = vlad.reshape(-1, cluster_size * F)                                   # :297 index f*K + c
    vlad = l2_normalize(vlad, 1)                                                # :298
    inter = {"vlad_flat": vlad.copy()} if return_intermediate else None
    Wh = V[vs + "hidden1_weights"]
    if pooling == "G_VLAD":
        vlad = vlad.reshape(-1, cluster_size * F // groups)                     # :302
    vlad = np.matmul(vlad, Wh)                                                  # :320 / :204
    vlad = fused_batch_norm_inference(vlad, V[vs + "bn/beta"], V[vs + "bn/gamma"],
                                      V[vs + "bn/moving_mean"], V[vs + "bn/moving_variance"])   # :321 / :207
    if pooling == "G_VLAD":
        vlad = vlad.reshape(-1, groups, output_dim)                             # :324
        vlad = np.sum(vlad, axis=-2, dtype=F32)                                 # :326
    if gating:
        vlad = context_gating(vlad, V, vs)                                      # :328-329
    if return_intermediate:
        return vlad, inter
    return vlad


# ------------------------------------------------------------------------------------------------
# models/*.py
# ------------------------------------------------------------------------------------------------
_ARCH = {  # backbone scope, blocks, head
    "epc-net": ("fastdgcnn", 4, "gvlad"),
    "kd_epc-net": ("fastdgcnn", 4, "gvlad"),
    "epc-net-l": ("fastdgcnn", 2, "maxfc"),
    "kd_epc-net-l": ("BACKBONE", 2, "maxfc"),
}


def forward(arch, point_cloud, V, params, scope="query_triplets", mask=None, arith="muladd",
            return_intermediate=False):
    """MODEL.forward(point_cloud, is_training=False, params=params).

    epc-net: models/epc-net.py:29-157; epc-net-l: models/epc-net-l.py:29-102; kd variants return
    ``(l2norm(per-point 1024 features) (B*N,1024), output)`` -- models/kd_epc-net.py:157-158,
    models/kd_epc-net-l.py:102.

    point_cloud: (Bq, P, N, 3) fp32.  Returns (Bq, P, OUTPUT_DIM).
    ``mask`` lets a caller inject the (B,N,N) 0/1 matrix computed by the C oracle (either arithmetic).
    """
    bscope, nblk, head = _ARCH[arch]
    pc = np.asarray(point_cloud, dtype=F32)
    Bq, P, N, dim = pc.shape
    k = params.get("KNN", 20)
    out_dim = params.get("FEATURE_OUTPUT_DIM", 256)
    pc = pc.reshape(Bq * P, N, dim)                                             # epc-net.py:41
    B = Bq * P
    if mask is None:
        mask = pairwise_distance_mask(pc, k=k, arith=arith)                     # :63
    inter = {}
    bs = "%s/%s/" % (scope, bscope)
    x_prev = pc
    feats = []
    for b in range(1, nblk + 1):
        x = conv1d(x_prev, V, bs + "conv%d" % b)                                # :66-69
        m = np.matmul(mask, x)                                                  # :70 dense BxNxN @ BxNx64
        m = m / F32(float(k))                                                   # :71
        t = m - x                                                               # :72
        t = conv1d(t, V, bs + "conv%d_a" % b)                                   # :73-76
        t = conv1d(t, V, bs + "conv%d_b" % b)                                   # :77-80
        xb = t + m                                                              # :81
        feats.append(xb)
        x_prev = xb
    xc = np.concatenate(feats, axis=-1)                                         # :134
    H = conv1d(xc, V, bs + "conv5")                                             # :136-139  (B,N,1024)
    if return_intermediate:
        inter["concat"] = xc
        inter["conv5"] = H
    vs = scope + "/VLAD/"
    if head == "gvlad":
        net = H.reshape(-1, 1024)                                               # :146
        net = l2_normalize(net, 1)                                              # :147
        out = vlad_forward(net, V, vs, N, params.get("CLUSTER_SIZE", 64), out_dim,
                           params.get("GROUPS", 4))                            # :141-149
    else:
        g = np.max(H, axis=1)                                                   # epc-net-l.py:91 max_pool2d [N,1]
        out = fully_connected(g, V, vs + "fc1")                                 # :95 (default relu!)
    out = l2_normalize(out, 1)                                                  # :153 / -l :98
    out = out.reshape(Bq, P, out_dim)                                           # :155
    if arch.startswith("kd_"):
        feat = l2_normalize(H.reshape(-1, 1024), 1)
        res = (feat, out)
    else:
        res = out
    if return_intermediate:
        return res, inter
    return res


def forward_f64(arch, point_cloud, V, params, scope="query_triplets", mask=None, arith="muladd"):
    """fp64 shadow of ``forward``: the same graph evaluated in float64 on the SAME neighbour mask (the mask is
    a discrete decision and stays pinned to the canonical fp32 arithmetic).  It bounds the fp32 restatement's own rounding
    error, so that a CUDA-vs-oracle difference can be split into "oracle noise" and "kernel error"."""
    global F32
    pc32 = np.asarray(point_cloud, dtype=np.float32)
    Bq, P, N, dim = pc32.shape
    if mask is None:
        mask = pairwise_distance_mask(pc32.reshape(Bq * P, N, dim), k=params.get("KNN", 20), arith=arith)
    saved = F32
    F32 = np.float64
    try:
        V64 = {k: np.asarray(v, dtype=np.float64) for k, v in V.items()}
        return forward(arch, pc32.astype(np.float64), V64, params, scope=scope, mask=np.asarray(mask, dtype=np.float64))
    finally:
        F32 = saved


# ------------------------------------------------------------------------------------------------
# evaluate.py
# ------------------------------------------------------------------------------------------------
def get_latent_vectors(arch, V, params, data, batch_num_queries=1, positives=0, negatives=0, **kw):
    """evaluate.get_latent_vectors -- evaluate.py:351-452.  ``data``: (n, N, 3)."""
    data = np.asarray(data, dtype=F32)
    n = data.shape[0]
    N, dim = data.shape[1], data.shape[2]
    batch_num = batch_num_queries * (1 + positives + negatives)                 # :355
    outs = []
    for q in range(n // batch_num):                                             # :357
        chunk = data[q * batch_num:(q + 1) * batch_num]
        q1 = chunk[0:batch_num_queries][:, None]                                # :368-369
        q2 = chunk[batch_num_queries:batch_num_queries * (positives + 1)].reshape(batch_num_queries, positives, N, dim)
        q3 = chunk[batch_num_queries * (positives + 1):].reshape(batch_num_queries, negatives, N, dim)
        vecs = np.concatenate([q1, q2, q3], axis=1)                             # evaluate.py:249
        o = forward(arch, vecs, V, params, **kw)
        if isinstance(o, tuple):
            o = o[1]
        o1, o2, o3 = o[:, :1], o[:, 1:1 + positives], o[:, 1 + positives:]      # :251
        outs.append(np.vstack([o1.reshape(-1, o.shape[-1]), o2.reshape(-1, o.shape[-1]),
                               o3.reshape(-1, o.shape[-1])]))                   # :402-407
    q_output = np.concatenate(outs, 0) if outs else np.zeros((0, params.get("FEATURE_OUTPUT_DIM", 256)), F32)
    for idx in range(n // batch_num * batch_num, n):                            # :415 tail: zero "fake" clouds
        queries = data[idx][None, None]
        fake = np.zeros((batch_num_queries - 1, 1, N, dim), F32)                # :425-430
        qq = np.vstack([queries, fake])
        vecs = np.concatenate([qq, np.zeros((batch_num_queries, positives + negatives, N, dim), F32)], axis=1)
        o = forward(arch, vecs, V, params, **kw)
        if isinstance(o, tuple):
            o = o[1]
        q_output = np.vstack([q_output, o[0, 0][None]])                         # :438-444
    return q_output

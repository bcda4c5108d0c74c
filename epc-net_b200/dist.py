"""Multi-GPU sharding of the path.  One process per GPU, torch.distributed plumbing.

  * embedding: independent clouds -> contiguous split of the batch, NO data-path collective
    (optional all_gather of the 1 KB descriptors when one consumer needs them all);
  * retrieval: database rows sharded D/G per rank, queries replicated; each rank finds its local top-k with
    GLOBAL row ids, one all_gather of (dist fp64, idx int64) [Q,k] per rank, then a (distance, index) merge --
    the result is independent of the shard count.

The reference has no distributed code; this is the B200-native scale-out of evaluate.get_latent_vectors /
get_recall.  The compute callables are parameters so the host logic can be exercised with gloo on CPU.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int):
    """Contiguous split of n items: first (n % world) ranks get one extra."""
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def embed_sharded(embed_fn, clouds, rank: int = None, world: int = None, gather: bool = False, group=None):
    """Embed this rank's slice of ``clouds`` (host array [n,N,3]) with ``embed_fn(host_slice) -> [m,D]`` host array.
    Returns (descriptors of the local slice, (start, end)); with gather=True every rank gets all n descriptors."""
    rank = dist.get_rank(group) if rank is None else rank
    world = dist.get_world_size(group) if world is None else world
    s, e = shard_range(len(clouds), rank, world)
    local = np.asarray(embed_fn(clouds[s:e]), dtype=np.float32)
    if not gather:
        return local, (s, e)
    D = local.shape[1]
    counts = [shard_range(len(clouds), r, world) for r in range(world)]
    mx = max(b - a for a, b in counts)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    buf = torch.zeros((mx, D), dtype=torch.float32, device=dev)
    buf[:e - s] = torch.from_numpy(local).to(dev)
    outs = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(outs, buf, group=group)
    full = np.concatenate([outs[r][:counts[r][1] - counts[r][0]].cpu().numpy() for r in range(world)], 0)
    return full, (s, e)


def retrieve_sharded(local_topk_fn, merge_fn, db_local, id_offset: int, queries, k: int, group=None):
    """db_local: this rank's rows (global ids id_offset..); queries: replicated [Q,dim].
    local_topk_fn(db_local, queries, k, id_offset) -> (dist [Q,k] float64, idx [Q,k] int64) torch tensors;
    merge_fn(dist [R,Q,k], idx [R,Q,k]) -> (dist [Q,k], idx [Q,k]).
    Rows with idx < 0 are padding (a shard smaller than k).
    The (distance, index) lists travel as ONE packed buffer through a single all-gather (the int64 ids ride as their
    float64 bit patterns)."""
    world = dist.get_world_size(group)
    d, i = local_topk_fn(db_local, queries, k, id_offset)
    Q = d.shape[0]
    packed = torch.empty((2, Q, k), dtype=torch.float64, device=d.device)
    packed[0].copy_(d)
    packed[1].copy_(i.view(torch.float64))
    g = torch.empty((world, 2, Q, k), dtype=torch.float64, device=d.device)
    if d.is_cuda:
        dist.all_gather_into_tensor(g, packed, group=group)
    else:
        dist.all_gather(list(g.unbind(0)), packed, group=group)
    return merge_fn(g[:, 0], g[:, 1].view(torch.int64))


def cuda_local_topk(db_local, queries, k, id_offset):
    from . import evaluate
    return evaluate.retrieve_topk(db_local, queries, k, id_offset)


def _merge_strided(g, R, Q, k):
    """g: [R, 2, Q, k] float64, g[r, 0] distances, g[r, 1] int64 row ids (bit patterns) -> merged (dist, idx) [Q, k]."""
    from . import _lib
    from .engine import _ptr, _stream
    od = torch.empty((Q, k), dtype=torch.float64, device=g.device)
    oi = torch.empty((Q, k), dtype=torch.int64, device=g.device)
    assert g.is_contiguous() and g.dtype == torch.float64 and tuple(g.shape) == (R, 2, Q, k)
    with torch.cuda.device(g.device):
        _lib.check(_lib.load().epc_merge_topk_strided(_ptr(g[0, 0]), _ptr(g[0, 1]), 2 * Q * k, R, Q, k, _ptr(od), _ptr(oi), _stream()))
    return od, oi


def cuda_merge(gd, gi):
    from . import _lib
    from .engine import _ptr, _stream
    R, Q, k = gd.shape
    if (not gd.is_contiguous() and gd.stride() == gi.stride() and gd.stride(1) == k and gd.stride(2) == 1
            and gd.stride(0) == 2 * Q * k and gi.data_ptr() == gd.data_ptr() + Q * k * 8):
        # the two halves of one packed all-gather buffer: merge in place, no repacking copy
        od = torch.empty((Q, k), dtype=torch.float64, device=gd.device)
        oi = torch.empty((Q, k), dtype=torch.int64, device=gd.device)
        with torch.cuda.device(gd.device):
            _lib.check(_lib.load().epc_merge_topk_strided(_ptr(gd), _ptr(gi), 2 * Q * k, R, Q, k, _ptr(od), _ptr(oi), _stream()))
        return od, oi
    od = torch.empty((Q, k), dtype=torch.float64, device=gd.device)
    oi = torch.empty((Q, k), dtype=torch.int64, device=gd.device)
    with torch.cuda.device(gd.device):
        _lib.check(_lib.load().epc_merge_topk(_ptr(gd.contiguous()), _ptr(gi.contiguous()), R, Q, k, _ptr(od), _ptr(oi), _stream()))
    return od, oi


class ShardedRetrieval:
    """evaluate.get_recall's ``KDTree(database_output)`` (evaluate.py:463) over a database whose rows are sharded across the
    ranks of ``group``: the rank's shard is prepared once (evaluate.RetrievalIndex, global row ids), every ``query`` finds the
    local top-k straight into one packed (dist | idx) buffer, all-gathers it in a single NCCL call and merges by
    (distance, index) -- the result does not depend on the shard count.  ``database_output``: the FULL [D, dim] host array
    (each rank keeps only its slice on the device) or, with ``local=True``, this rank's rows."""

    def __init__(self, database_output, group=None, local=False, id_offset=0):
        from . import evaluate
        self.group = group
        self.world = dist.get_world_size(group)
        rank = dist.get_rank(group)
        if local:
            rows, off = database_output, int(id_offset)
        else:
            s, e = shard_range(len(database_output), rank, self.world)
            rows, off = database_output[s:e], s
        self.index = evaluate.RetrievalIndex(rows, id_offset=off)

    def query(self, queries_output, k):
        from .engine import as_cuda_f32
        q = as_cuda_f32(queries_output, "queries_output")
        Q, k = q.shape[0], int(k)
        packed = torch.empty((2, Q, k), dtype=torch.float64, device=q.device)
        self.index.query(q, k, out=(packed[0], packed[1].view(torch.int64)))
        g = torch.empty((self.world, 2, Q, k), dtype=torch.float64, device=q.device)
        dist.all_gather_into_tensor(g, packed, group=self.group)
        return _merge_strided(g, self.world, Q, k)

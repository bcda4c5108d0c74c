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


def make_grid_groups(db_shards: int):
    """Process groups of a (query groups x database shards) layout of the default world: rank = qg * db_shards + ds.
    Returns (ds, qg, q_groups, shard_group, peer_group): shard_group = the db_shards ranks that hold the shards of one query
    group's database copy (their candidate lists are merged); peer_group = the ranks with the same shard index across the
    query groups (they exchange the merged results of their query slices).  Collective: every rank must call it."""
    world, rank = dist.get_world_size(), dist.get_rank()
    if db_shards < 1 or world % db_shards != 0:
        raise ValueError("db_shards=%d must divide the world size %d" % (db_shards, world))
    q_groups = world // db_shards
    ds, qg = rank % db_shards, rank // db_shards
    shard_group = peer_group = None
    if q_groups > 1:
        for g in range(q_groups):
            h = dist.new_group([g * db_shards + d for d in range(db_shards)])
            if g == qg:
                shard_group = h
        for d in range(db_shards):
            h = dist.new_group([g * db_shards + d for g in range(q_groups)])
            if d == ds:
                peer_group = h
    return ds, qg, q_groups, shard_group, peer_group


def _all_gather_packed(packed, n, group):
    g = torch.empty((n,) + tuple(packed.shape), dtype=packed.dtype, device=packed.device)
    if packed.is_cuda:
        dist.all_gather_into_tensor(g, packed, group=group)
    else:
        dist.all_gather(list(g.unbind(0)), packed, group=group)
    return g


def retrieve_sharded_2d(local_topk_fn, merge_fn, db_local, id_offset: int, queries, k: int, grid):
    """Database sharded ``db_shards`` ways AND queries split over ``q_groups`` groups (grid = make_grid_groups(db_shards)):
    a rank scores its query slice against its database shard, the db_shards lists of a slice are all-gathered and merged
    inside the shard group, and the merged slices are all-gathered across the query groups.  Every rank returns the full
    (dist [Q,k], idx [Q,k]).  With q_groups == 1 this is retrieve_sharded.  Per-rank work is 1/world of the scoring AND
    1/q_groups of the per-query work (re-rank, selection, merge) that pure database sharding repeats on every rank."""
    ds, qg, q_groups, shard_group, peer_group = grid
    db_shards = dist.get_world_size() // q_groups
    Q = queries.shape[0]
    s, e = shard_range(Q, qg, q_groups)
    d, i = local_topk_fn(db_local, queries[s:e], k, id_offset)
    nq = e - s
    packed = torch.empty((2, nq, k), dtype=torch.float64, device=d.device)
    packed[0].copy_(d)
    packed[1].copy_(i.view(torch.float64))
    g = _all_gather_packed(packed, db_shards, shard_group)
    md, mi = merge_fn(g[:, 0], g[:, 1].view(torch.int64))
    if q_groups == 1:
        return md, mi
    nmax = -(-Q // q_groups)
    mine = torch.zeros((2, nmax, k), dtype=torch.float64, device=md.device)
    mine[0, :nq].copy_(md)
    mine[1, :nq].copy_(mi.view(torch.float64))
    allq = _all_gather_packed(mine, q_groups, peer_group)
    od = torch.empty((Q, k), dtype=torch.float64, device=md.device)
    oi = torch.empty((Q, k), dtype=torch.int64, device=md.device)
    for gq in range(q_groups):
        a, b = shard_range(Q, gq, q_groups)
        od[a:b].copy_(allq[gq, 0, :b - a])
        oi[a:b].copy_(allq[gq, 1, :b - a].view(torch.int64))
    return od, oi


class ShardedRetrieval:
    """evaluate.get_recall's ``KDTree(database_output)`` (evaluate.py:463) over a database whose rows are sharded across the
    ranks: the rank's shard is prepared once (evaluate.RetrievalIndex, global row ids), every ``query`` finds the local
    top-k straight into one packed (dist | idx) buffer, all-gathers it in a single NCCL call and merges by (distance, index)
    -- the result does not depend on the shard count.  ``database_output``: the FULL [D, dim] host array (each rank keeps
    only its slice on the device) or, with ``local=True``, this rank's rows.
    ``db_shards`` < world size: the ranks form (world / db_shards) query groups, each holding one sharded copy of the database
    and answering 1/groups of the queries (retrieve_sharded_2d); the default is one group = the database sharded over every
    rank (BASELINE.json configs[4])."""

    def __init__(self, database_output, group=None, local=False, id_offset=0, db_shards=None):
        from . import evaluate
        self.group = group
        self.world = dist.get_world_size(group)
        rank = dist.get_rank(group)
        self.db_shards = int(db_shards) if db_shards else self.world
        self.grid = None
        if self.db_shards != self.world:
            if group is not None or local:
                raise ValueError("db_shards < world size needs the default process group and the full database array")
            self.grid = make_grid_groups(self.db_shards)
            rank = self.grid[0]
        if local:
            rows, off = database_output, int(id_offset)
        else:
            s, e = shard_range(len(database_output), rank, self.db_shards)
            rows, off = database_output[s:e], s
        self.index = evaluate.RetrievalIndex(rows, id_offset=off)

    def query(self, queries_output, k):
        from .engine import as_cuda_f32
        q = as_cuda_f32(queries_output, "queries_output")
        Q, k = q.shape[0], int(k)
        if self.grid is not None:
            return retrieve_sharded_2d(lambda _db, qq, kk, _off: self.index.query(qq, kk), cuda_merge, None, 0, q, k, self.grid)
        packed = torch.empty((2, Q, k), dtype=torch.float64, device=q.device)
        self.index.query(q, k, out=(packed[0], packed[1].view(torch.int64)))
        g = torch.empty((self.world, 2, Q, k), dtype=torch.float64, device=q.device)
        dist.all_gather_into_tensor(g, packed, group=self.group)
        return _merge_strided(g, self.world, Q, k)

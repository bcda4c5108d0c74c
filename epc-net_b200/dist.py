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
    Rows with idx < 0 are padding (a shard smaller than k)."""
    world = dist.get_world_size(group)
    d, i = local_topk_fn(db_local, queries, k, id_offset)
    d = d.contiguous()
    i = i.contiguous()
    gd = torch.empty((world,) + tuple(d.shape), dtype=d.dtype, device=d.device)
    gi = torch.empty((world,) + tuple(i.shape), dtype=i.dtype, device=i.device)
    dist.all_gather_into_tensor(gd, d, group=group) if d.is_cuda else dist.all_gather(list(gd.unbind(0)), d, group=group)
    dist.all_gather_into_tensor(gi, i, group=group) if i.is_cuda else dist.all_gather(list(gi.unbind(0)), i, group=group)
    return merge_fn(gd, gi)


def cuda_local_topk(db_local, queries, k, id_offset):
    from . import evaluate
    return evaluate.retrieve_topk(db_local, queries, k, id_offset)


def cuda_merge(gd, gi):
    import ctypes
    from . import _lib
    from .engine import _ptr, _stream
    R, Q, k = gd.shape
    od = torch.empty((Q, k), dtype=torch.float64, device=gd.device)
    oi = torch.empty((Q, k), dtype=torch.int64, device=gd.device)
    with torch.cuda.device(gd.device):
        _lib.check(_lib.load().epc_merge_topk(_ptr(gd.contiguous()), _ptr(gi.contiguous()), R, Q, k, _ptr(od), _ptr(oi), _stream()))
    return od, oi

"""Multi-GPU plumbing (torch.distributed only — not the product): queries shard across ranks, the index is replicated.

  * replicate_index: rank 0 holds the built index; every other rank allocates a same-shape replica and the arrays
    travel by one `dist.broadcast` each over NCCL/NVLink (the north-star's "NCCL broadcast at load only").
  * shard_queries / search_sharded: contiguous query shards, no data-path collective; per-rank results are gathered
    on the host of rank 0 (gather_object), exactly as N independent backends would return them.
"""
from __future__ import annotations

import numpy as np


class _DevArray:
    """Zero-copy view of a raw device pointer for torch (CUDA array interface v3)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False),
                                         "version": 3}


def replicate_index(index, rank: int, device: int, src: int = 0):
    """Collective. `index` is the built Index on rank `src` and None elsewhere. Returns this rank's Index."""
    import torch
    import torch.distributed as dist
    from .bm25x import Index, IndexLayout, N_ARRAYS

    meta = [None]
    if rank == src:
        lay = index.layout()
        meta[0] = {f: getattr(lay, f) for f in ("n_docs", "n_terms", "n_postings", "n_postings_padded", "n_blocks",
                                                "sum_doc_len", "k1", "b", "avgdl")}
        meta[0]["bytes"] = [int(x) for x in lay.bytes]
    dist.broadcast_object_list(meta, src=src)
    if rank != src:
        like = IndexLayout()
        for f, v in meta[0].items():
            if f != "bytes":
                setattr(like, f, v)
        index = Index.alloc_replica(like, device)
        lay = index.layout()
        assert [int(x) for x in lay.bytes] == meta[0]["bytes"], "replica layout mismatch"
    for i in range(N_ARRAYS):
        t = torch.as_tensor(_DevArray(lay.dev_ptr[i], lay.bytes[i]), device=torch.device("cuda", device))
        dist.broadcast(t, src=src)
    torch.cuda.synchronize()
    if rank != src:
        index.finalize_replica()
    return index


def shard_bounds(nq: int, rank: int, world: int):
    """Contiguous shard [lo, hi) of nq queries for `rank` (sizes differ by at most one)."""
    base, rem = divmod(nq, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_queries(q_off, q_terms, rank: int, world: int):
    q_off = np.asarray(q_off, dtype=np.uint32)
    lo, hi = shard_bounds(len(q_off) - 1, rank, world)
    sub_off = (q_off[lo:hi + 1] - q_off[lo]).astype(np.uint32)
    return sub_off, np.asarray(q_terms, dtype=np.uint32)[q_off[lo]:q_off[hi]], lo, hi


def gather_rows(parts: dict, n_total: int, rank: int, world: int, dst: int = 0):
    """Collective.  parts: {name: torch tensor [n_local, ...]} holding this rank's rows of a contiguous row sharding
    (shard_bounds) — CUDA tensors under NCCL (the rows travel GPU → GPU over NVLink), CPU tensors under gloo.
    Returns {name: tensor [n_total, ...]} on `dst` (same device as the inputs), None elsewhere.  Shards differ by at
    most one row: every rank contributes max-shard rows (one padded row at most) to ONE dist.gather per array."""
    import torch
    import torch.distributed as dist

    sizes = [shard_bounds(n_total, r, world)[1] - shard_bounds(n_total, r, world)[0] for r in range(world)]
    mx = max(sizes) if sizes else 0
    out = {}
    for name, t in parts.items():
        if t.shape[0] != mx:   # pad to the common shard size
            pad = torch.zeros((mx - t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
            t = torch.cat([t, pad], dim=0)
        t = t.contiguous()
        if world == 1:
            out[name] = t[:n_total]
            continue
        bufs = [torch.empty_like(t) for _ in range(world)] if rank == dst else None
        dist.gather(t, bufs, dst=dst)
        if rank == dst:
            out[name] = torch.cat([b[:sizes[r]] for r, b in enumerate(bufs)], dim=0)
    return out if rank == dst else None


def search_sharded(search_fn, q_off, q_terms, k: int, rank: int, world: int, dst: int = 0):
    """Collective. search_fn(q_off, q_terms, k) -> {"doc": [n,k], "score": [n,k], "n": [n], ...} on this rank's shard.
    Returns the concatenated result (query order preserved) on rank `dst`, None elsewhere."""
    import torch.distributed as dist

    sub_off, sub_terms, lo, hi = shard_queries(q_off, q_terms, rank, world)
    res = search_fn(sub_off, sub_terms, k)
    part = {key: np.asarray(v) for key, v in res.items() if isinstance(v, np.ndarray)}
    parts = [None] * world if rank == dst else None
    if world > 1:
        dist.gather_object(part, parts, dst=dst)
    else:
        parts = [part]
    if rank != dst:
        return None
    return {key: np.concatenate([p[key] for p in parts], axis=0) for key in parts[0]}

"""Batch sharding across GPUs: independent problems, no collective.

Problem j of the global batch goes to rank j mod world (interleaved, so the spread of
iteration counts evens out); every rank runs its own persistent grid and work queue.
The only communication is the final gather of results, which is host-side plumbing
(torch.distributed all_gather over gloo/NCCL), never part of the data path.
"""
import numpy as np


def shard_columns(batch, rank, world):
    """Global problem indices owned by `rank`."""
    return np.arange(rank, batch, world, dtype=np.int64)


def scatter(Theta, rank, world):
    """This rank's columns of the n×batch start matrix (column-major copy)."""
    return np.asfortranarray(Theta[:, rank::world])


def gather(parts, batch):
    """Inverse of scatter for per-problem results.  parts[r] has the problems of rank r
    along its LAST axis; returns the array in global problem order."""
    world = len(parts)
    first = np.asarray(parts[0])
    out = np.empty(first.shape[:-1] + (batch,), dtype=first.dtype)
    for r, p in enumerate(parts):
        out[..., r::world] = p
    return out


def all_gather_results(local, batch, rank, world, dist):
    """Gather a dict of per-problem arrays (problem axis last) from every rank with
    torch.distributed (any backend).  Ragged shards are padded to the largest."""
    import torch
    nmax = (batch + world - 1) // world
    out = {}
    for k, a in local.items():
        a = np.ascontiguousarray(a)
        pad = np.zeros(a.shape[:-1] + (nmax,), dtype=a.dtype)
        pad[..., :a.shape[-1]] = a
        t = torch.from_numpy(pad)
        bufs = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(bufs, t)
        parts = [b.numpy()[..., :len(shard_columns(batch, r, world))] for r, b in enumerate(bufs)]
        out[k] = gather(parts, batch)
    return out


ROWS_PER_TILE = 512  # kRowsPerCta of csrc/qn_stream.cu: 256 threads x 2 rows


def row_tile_partition(n, nshard):
    """The row-sharded large-n mode's split of B⁻¹ (qn_bfgs_optimize_rowsharded_f64, stream_setup in
    csrc/qn_stream.cu): shard g of G owns the 512-row tiles [T*g//G, T*(g+1)//G) with
    T = ceil(n/512) and G = min(nshard, T).  Returns [(row_begin, row_end)] per shard."""
    tiles = (n + ROWS_PER_TILE - 1) // ROWS_PER_TILE
    G = min(int(nshard), tiles)
    out = []
    for g in range(G):
        t0, t1 = tiles * g // G, tiles * (g + 1) // G
        out.append((t0 * ROWS_PER_TILE, min(n, t1 * ROWS_PER_TILE)))
    return out

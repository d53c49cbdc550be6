"""Host-side plan of the row-sharded large-n mode on CPU: the row-tile split
mirrored in sharding.row_tile_partition, and — with world_size-2 gloo — the property the mode rests
on: a row sum depends only on its own row of B⁻¹, so all-gathering per-shard slices of t = B⁻¹y
reproduces the single-rank vector bit for bit (no all-reduce, hence no change of summation order)."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT


def test_row_tile_partition_covers_rows_once():
    sys.path.insert(0, ROOT)
    from qn_b200 import sharding
    for n in (2, 511, 512, 513, 1100, 2049, 8192, 16384, 10000):
        for G in (1, 2, 3, 4, 8):
            parts = sharding.row_tile_partition(n, G)
            assert 1 <= len(parts) <= G
            assert parts[0][0] == 0 and parts[-1][1] == n
            for (a0, a1), (b0, b1) in zip(parts, parts[1:]):
                assert a1 == b0 and a0 < a1
            for r0, r1 in parts:
                assert r0 % sharding.ROWS_PER_TILE == 0 and r1 > r0
    assert sharding.row_tile_partition(8192, 8) == [(1024 * g, 1024 * (g + 1)) for g in range(8)]
    assert len(sharding.row_tile_partition(600, 8)) == 2  # a shard owns at least one 512-row tile


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rowsum(H, y):
    """row sums in one fixed order (ascending columns, one fma-free chain per row)"""
    t = np.zeros(H.shape[0])
    for r in range(H.shape[1]):
        t = t + H[:, r] * y[r]
    return t


def _worker(rank, world, port, n, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    from qn_b200 import sharding
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.Generator(np.random.Philox(key=77))
    H = rng.standard_normal((n, n))
    y = rng.standard_normal(n)
    r0, r1 = sharding.row_tile_partition(n, world)[rank]
    mine = _rowsum(H[r0:r1], y)                      # this shard's slice of t = B⁻¹y
    sizes = [b - a for a, b in sharding.row_tile_partition(n, world)]
    bufs = [torch.zeros(s, dtype=torch.float64) for s in sizes]
    dist.all_gather(bufs, torch.from_numpy(mine)) if len(set(sizes)) == 1 else \
        [dist.broadcast(bufs[g] if g != rank else bufs[g].copy_(torch.from_numpy(mine)), src=g) for g in range(world)]
    t = torch.cat(bufs).numpy()
    # replicated reduction: every rank forms yᵀB⁻¹y from the SAME gathered vector in the SAME order
    yHy = float(np.sum(t * y))
    both = torch.tensor([yHy], dtype=torch.float64)
    lst = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(lst, both)
    assert all(v.item() == yHy for v in lst)
    if rank == 0:
        np.save(os.path.join(out_dir, "t.npy"), t)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [1024, 1500])
def test_two_rank_row_slices_gather_to_the_single_rank_vector(tmp_path, n):
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.Generator(np.random.Philox(key=77))
    H = rng.standard_normal((n, n))
    y = rng.standard_normal(n)
    assert np.array_equal(np.load(os.path.join(str(tmp_path), "t.npy")), _rowsum(H, y))

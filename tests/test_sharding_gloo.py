"""The N>1 path on CPU: world_size-2 `gloo`, one process per (pretend) GPU.

The data path has no collective (independent problems); what is tested here is the
host-side plan bench.py and the package use: interleaved scatter of the starts, a
per-rank solve, and the gather of per-problem results in global order.  The per-rank
solver is the CPU oracle standing in for the kernel (the GPU box has the real one).
"""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT, starts


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, batch, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    import qn_b200
    from qn_b200 import sharding
    from oracle import binding as ob
    dist.init_process_group("gloo", rank=rank, world_size=world)
    Theta = starts(n, batch)
    mine = sharding.scatter(Theta, rank, world)
    r = ob.optimize_batch(ob.rosenbrock(), mine, nthreads=1)
    local = dict(x=r["x"], ell=r["ell"], iters=r["iters"], status=r["status"])
    full = sharding.all_gather_results(local, batch, rank, world, dist)
    # timing plumbing used by bench.py: max over ranks
    import torch
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert t.item() == world
    dist.barrier()
    if rank == 0:
        np.savez(os.path.join(out_dir, "gathered.npz"), **full)
    dist.destroy_process_group()


@pytest.mark.parametrize("batch", [64, 37])
def test_two_rank_shard_and_gather(tmp_path, ob, batch):
    import torch.multiprocessing as mp
    n, world = 12, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n, batch, str(tmp_path)), nprocs=world, join=True)
    got = np.load(os.path.join(str(tmp_path), "gathered.npz"))
    ref = ob.optimize_batch(ob.rosenbrock(), starts(n, batch), nthreads=1)
    for k in ("x", "ell", "iters", "status"):
        assert np.array_equal(got[k], ref[k], equal_nan=True), k

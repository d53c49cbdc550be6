"""Does the device-resident batch also gain from being launched as R ranges on R streams?

    python scripts/exp_device_ranges.py

n = 60 Rosenbrock, 2^16 starts resident in HBM; CUDA events on a parent stream around: one launch
(the device entry as bench.py's `value` leg calls it), and R = 2/4/8/16 launches of batch/R problems
forked onto R streams and joined back (events), with and without an L2 flush between repetitions."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import qn_b200 as qn
from bench import global_starts

n, batch = 60, 1 << 16
eng = qn.Engine(0)
obj = qn.Rosenbrock()
dev = torch.device("cuda:0")
th = torch.from_numpy(np.ascontiguousarray(global_starts(batch, n).T)).to(dev)  # [batch][n]
x = torch.empty_like(th); g = torch.empty_like(th)
ell = torch.empty(batch, dtype=torch.float64, device=dev)
it = torch.empty(batch, dtype=torch.int32, device=dev); fe = torch.empty_like(it); st = torch.empty_like(it)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
parent = torch.cuda.Stream()
streams = [torch.cuda.Stream() for _ in range(16)]


def launch(c0, cnt, s):
    eng.optimize_batched_device(obj, n, cnt, th[c0:].data_ptr(), x_ptr=x[c0:].data_ptr(), g_ptr=g[c0:].data_ptr(),
                                ell_ptr=ell[c0:].data_ptr(), iters_ptr=it[c0:].data_ptr(), fevals_ptr=fe[c0:].data_ptr(),
                                status_ptr=st[c0:].data_ptr(), stream=s.cuda_stream)


def step(R):
    if R < 0:  # -R ranges one after the other on the parent stream: no overlap, R tails
        R = -R
        for r in range(R):
            c0 = batch * r // R
            launch(c0, batch * (r + 1) // R - c0, parent)
        return
    if R == 1:
        launch(0, batch, parent)
        return
    fork = torch.cuda.Event(); fork.record(parent)
    for r in range(R):
        streams[r].wait_event(fork)
        c0 = batch * r // R
        launch(c0, batch * (r + 1) // R - c0, streams[r])
        e = torch.cuda.Event(); e.record(streams[r]); parent.wait_event(e)


ref = None
for parity in (-1, 0, 1, 2, 4, 8, 1 << 20):
    eng.set_option("lead_parity", parity)
    for R in (1, 4, -4):
        times = []
        for rep in range(7):
            with torch.cuda.stream(parent):
                flush.fill_(rep)
            a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
            a.record(parent); step(R); b.record(parent)
            torch.cuda.synchronize()
            if rep >= 2:
                times.append(a.elapsed_time(b))
        key = x.clone()
        if ref is None:
            ref = key
        print("lead_parity %7d  ranges %2d: mean %.3f ms  min %.3f ms  same bits %s" % (
            parity, R, sum(times) / len(times), min(times), bool(torch.equal(key, ref))), flush=True)

"""Which warp of a leader/helper pair leads (option "lead_parity"), per n and per tile home.

    python scripts/exp_lead_parity.py

Rosenbrock, 2^15 starts resident in HBM, CUDA events around the device entry, best of 4:
lead_parity 0 (the same warp in every CTA), -1 (alternating per residency slot, the round-2 first
choice) and 1 (alternating per CTA); tile in registers (tile_tmem 0) and tensor memory (1)."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import qn_b200 as qn
from bench import global_starts

eng = qn.Engine(0)
obj = qn.Rosenbrock()
dev = torch.device("cuda:0")
stream = torch.cuda.Stream()
cases = [(n, 1 << 15) for n in (36, 40, 44, 48, 52, 56, 60, 64)] + [(n, b) for n in (40, 60) for b in (148, 296, 592, 1184, 2368, 4736)]
for n, batch in cases:
    th = torch.from_numpy(np.ascontiguousarray(global_starts(batch, n).T)).to(dev)
    x = torch.empty_like(th); g = torch.empty_like(th)
    ell = torch.empty(batch, dtype=torch.float64, device=dev)
    it = torch.empty(batch, dtype=torch.int32, device=dev); fe = torch.empty_like(it); st = torch.empty_like(it)
    ref = None
    row = []
    for tm in (0, 1):
        eng.set_option("tile_tmem", tm)
        for parity in (0, -1, 1):
            eng.set_option("lead_parity", parity)
            best = 1e30
            for rep in range(5):
                a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
                a.record(stream)
                eng.optimize_batched_device(obj, n, batch, th.data_ptr(), x_ptr=x.data_ptr(), g_ptr=g.data_ptr(),
                                            ell_ptr=ell.data_ptr(), iters_ptr=it.data_ptr(), fevals_ptr=fe.data_ptr(),
                                            status_ptr=st.data_ptr(), stream=stream.cuda_stream)
                b.record(stream); torch.cuda.synchronize()
                if rep:
                    best = min(best, a.elapsed_time(b))
            if ref is None:
                ref = x.clone()
            assert torch.equal(x, ref)
            row.append("tm%d/p%2d %7.3f" % (tm, parity, best))
    print("n=%2d batch=%5d iters %.0f | %s" % (n, batch, it.sum().item() / batch, " | ".join(row)), flush=True)

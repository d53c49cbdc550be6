"""Which warp leads in the one-CTA-per-problem tile kernels (option "lead_warp"), n = 96, 128, 160.

    python scripts/exp_lead_warp.py

Rosenbrock (and the dense Gaussian at n = 128) resident in HBM, CUDA events around the device entry,
best of 3, every warp of the CTA as the leader."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import qn_b200 as qn
from bench import global_starts
from problems import gaussian_problem

eng = qn.Engine(0)
dev = torch.device("cuda:0")
stream = torch.cuda.Stream()
mu, lam = gaussian_problem(128)
cases = [("rosenbrock", qn.Rosenbrock(), 96, 8192), ("rosenbrock", qn.Rosenbrock(), 128, 4096),
         ("rosenbrock", qn.Rosenbrock(), 160, 2048), ("gaussian", qn.DenseGaussian(mu, lam), 128, 32768)]
for name, obj, n, batch in cases:
    th = torch.from_numpy(np.ascontiguousarray(global_starts(batch, n).T)).to(dev)
    x = torch.empty_like(th); g = torch.empty_like(th)
    ell = torch.empty(batch, dtype=torch.float64, device=dev)
    it = torch.empty(batch, dtype=torch.int32, device=dev); fe = torch.empty_like(it); st = torch.empty_like(it)
    ref = None
    row = []
    for lw in range(n // 16):
        eng.set_option("lead_warp", lw)
        best = 1e30
        for rep in range(4):
            a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
            a.record(stream)
            eng.optimize_batched_device(obj, n, batch, th.data_ptr(), x_ptr=x.data_ptr(), g_ptr=g.data_ptr(),
                                        ell_ptr=ell.data_ptr(), iters_ptr=it.data_ptr(), fevals_ptr=fe.data_ptr(),
                                        status_ptr=st.data_ptr(), stream=stream.cuda_stream)
            b.record(stream); torch.cuda.synchronize()
            if rep:
                best = min(best, a.elapsed_time(b))
        if ref is None:
            ref = (x.clone(), it.clone())
        assert torch.equal(x, ref[0]) and torch.equal(it, ref[1])
        row.append("w%d %.2f" % (lw, best))
    print("%s n=%d batch=%d block=%d | %s" % (name, n, batch, eng.last_launch()["block"], " | ".join(row)), flush=True)

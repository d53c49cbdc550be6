"""Two launches of one resident kernel through the device entry (for ncu: profile the second).

    python scripts/one_launch.py N BATCH [option=value ...]        e.g.  60 16384 lead_parity=-1
"""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import qn_b200 as qn
from bench import global_starts

n, batch = int(sys.argv[1]), int(sys.argv[2])
eng = qn.Engine(0)
for kv in sys.argv[3:]:
    k, v = kv.split("=")
    eng.set_option(k, int(v))
dev = torch.device("cuda:0")
th = torch.from_numpy(np.ascontiguousarray(global_starts(batch, n).T)).to(dev)
x = torch.empty_like(th); g = torch.empty_like(th)
ell = torch.empty(batch, dtype=torch.float64, device=dev)
it = torch.empty(batch, dtype=torch.int32, device=dev); fe = torch.empty_like(it); st = torch.empty_like(it)
for rep in range(2):
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    eng.optimize_batched_device(qn.Rosenbrock(), n, batch, th.data_ptr(), x_ptr=x.data_ptr(), g_ptr=g.data_ptr(),
                                ell_ptr=ell.data_ptr(), iters_ptr=it.data_ptr(), fevals_ptr=fe.data_ptr(),
                                status_ptr=st.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
    b.record(); torch.cuda.synchronize()
print("n=%d batch=%d %s: %.3f ms, %d converged, launch %s" % (n, batch, sys.argv[3:], a.elapsed_time(b),
      int((st == 0).sum().item()), eng.last_launch()))

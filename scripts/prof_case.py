"""One batched device-resident solve of a chosen shape, for ncu captures and quick timings.

    python scripts/prof_case.py --n 128 --batch 148 --obj rosenbrock [--reps 2] [--smem] [--max-iter K]
"""
import argparse, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import qn_b200 as qn
from problems import gaussian_problem, logistic_problem

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=128)
ap.add_argument("--batch", type=int, default=148)
ap.add_argument("--obj", default="rosenbrock")
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--smem", action="store_true")
ap.add_argument("--max-iter", type=int, default=10000)
a = ap.parse_args()
eng = qn.Engine(0)
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream()
n, batch = a.n, a.batch
if a.obj == "rosenbrock":
    obj = qn.Rosenbrock()
elif a.obj == "gaussian":
    obj = qn.DenseGaussian(*gaussian_problem(n))
else:
    obj = qn.LogisticRegression(*logistic_problem(1024, n))
rng = np.random.Generator(np.random.Philox(key=20261018 + n))
th = torch.from_numpy(rng.standard_normal((batch, n))).to(dev)
x = torch.empty_like(th); g = torch.empty_like(th)
ell = torch.empty(batch, dtype=torch.float64, device=dev)
it = torch.empty(batch, dtype=torch.int32, device=dev); fe = torch.empty_like(it); st = torch.empty_like(it)
flags = qn._lib.FLAG_FORCE_SMEM if a.smem else 0
for r in range(a.reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    eng.optimize_batched_device(obj, n, batch, th.data_ptr(), x_ptr=x.data_ptr(), g_ptr=g.data_ptr(), ell_ptr=ell.data_ptr(),
                                iters_ptr=it.data_ptr(), fevals_ptr=fe.data_ptr(), status_ptr=st.data_ptr(),
                                stream=stream.cuda_stream, flags=flags, max_iter=a.max_iter)
    e1.record(stream); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    iters = int(it.sum())
    print("%s n=%d batch=%d: %.3f ms, %d iterations (%.1f per solve), %.3e it/s, fevals/iter %.2f, launch %s" % (
        a.obj, n, batch, ms, iters, iters / batch, iters / (ms * 1e-3), float(fe.sum()) / max(iters, 1), eng.last_launch()), flush=True)

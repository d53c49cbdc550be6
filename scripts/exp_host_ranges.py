"""End-to-end time of the host-pointer entry against option "host_chunks" (pinned host buffers).

    python scripts/exp_host_ranges.py [--gpus G]

n = 60 Rosenbrock, 2^16 starts (the headline batch): wall clock around qn_bfgs_optimize_batched_f64
(G = 1) or qn_bfgs_optimize_batched_multi_f64 (G > 1), best of 5, for 1 / 2 / 4 / 8 ranges; beside it
the device-resident time of the same batch (CUDA events around the device entry)."""
import argparse, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import qn_b200 as qn

ap = argparse.ArgumentParser()
ap.add_argument("--gpus", type=int, default=1)
ap.add_argument("--n", type=int, default=60)
ap.add_argument("--batch", type=int, default=1 << 16)
args = ap.parse_args()
G, n, batch = args.gpus, args.n, args.batch
engines = [qn.Engine(g) for g in range(G)]
rng = np.random.Generator(np.random.Philox(key=20261018))


def pinned(shape, dtype):
    return torch.empty(shape, dtype=dtype).pin_memory()


th = pinned((batch, n), torch.float64)
th.numpy()[...] = rng.standard_normal((batch, n))
Theta = th.numpy().T
out = dict(x=pinned((batch, n), torch.float64).numpy().T, g=pinned((batch, n), torch.float64).numpy().T,
           ell=pinned((batch,), torch.float64).numpy(), iters=pinned((batch,), torch.int32).numpy(),
           fevals=pinned((batch,), torch.int32).numpy(), status=pinned((batch,), torch.int32).numpy())
ref = None
for chunks in (1, 2, 4, 8, -1):
    for e in engines:
        e.set_option("host_chunks", chunks)
    best = 1e30
    for rep in range(6):
        t = time.perf_counter()
        r = qn.optimize_batched_multi(engines, qn.Rosenbrock(), Theta, out=out)
        dt = time.perf_counter() - t
        if rep:
            best = min(best, dt)
    key = (r["x"].copy(), r["iters"].copy())
    if ref is None:
        ref = key
    same = np.array_equal(key[0], ref[0]) and np.array_equal(key[1], ref[1])
    print("G=%d n=%d batch=%d host_chunks=%2d: %8.3f ms  %10.0f solves/s  same bits %s" % (
        G, n, batch, chunks, best * 1e3, int((r["status"] == 0).sum()) / best, same), flush=True)

if G == 1:
    import ctypes as C
    eng = engines[0]
    handle = eng.model_handle(qn.Rosenbrock())
    lsc = qn.BackTracking()._c()
    p = lambda a: a.ctypes.data
    for pageable in (False, True):
        if pageable:
            ThetaP = np.array(Theta, order="F")
            o = {k: np.empty_like(v) if k in ("ell", "iters", "fevals", "status") else np.empty((n, batch), order="F") for k, v in out.items()}
        else:
            ThetaP, o = Theta, out
        for chunks in (1, 2, 4, 8, -1):
            eng.set_option("host_chunks", chunks)
            best = 1e30
            for rep in range(6):
                t = time.perf_counter()
                qn._lib.check(eng._lib.qn_bfgs_optimize_batched_f64(
                    eng._ctx, handle, n, batch, p(ThetaP), C.byref(lsc), 1e-8, 10000, None, 0, 0,
                    p(o["x"]), p(o["g"]), p(o["ell"]), p(o["iters"]), p(o["fevals"]), p(o["status"]), 0, None, 0))
                dt = time.perf_counter() - t
                if rep:
                    best = min(best, dt)
            same = np.array_equal(o["x"], ref[0]) and np.array_equal(o["iters"], ref[1])
            print("host entry (%s buffers) host_chunks=%2d: %8.3f ms  %10.0f solves/s  same bits %s" % (
                "pageable" if pageable else "pinned", chunks, best * 1e3, int((o["status"] == 0).sum()) / best, same), flush=True)

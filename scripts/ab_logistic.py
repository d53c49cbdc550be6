"""A/B of two builds of libqn_cuda.so on the n = 128 logistic batch: time and a hash of every output.

    QN_LIBRARY_PATH=<lib> python scripts/ab_logistic.py [batch]

Run once per library; equal hashes = equal bits for every problem (x, g, ell, iters, fevals, status)."""
import hashlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import qn_b200 as qn
from problems import logistic_problem

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
X, y, l2 = logistic_problem(1024, 128)
obj = qn.LogisticRegression(X, y, l2)
rng = np.random.Generator(np.random.Philox(key=20261018 + 128))
Theta = np.asfortranarray(rng.standard_normal((batch, 128)).T)
eng = qn.Engine(0)
eng.optimize_batched(obj, Theta[:, :64])
best = 1e30
for _ in range(2):
    t = time.perf_counter(); r = eng.optimize_batched(obj, Theta); best = min(best, time.perf_counter() - t)
h = hashlib.sha256()
for k in ("x", "g", "ell", "iters", "fevals", "status"):
    h.update(np.ascontiguousarray(r[k]).tobytes())
conv = int((r["status"] == 0).sum())
print("%s: batch %d  %.2f ms  %.0f converged solves/s  conv %d  iters/solve %.1f  fevals/iter %.2f  hash %s" % (
    os.environ.get("QN_LIBRARY_PATH", "in-tree"), batch, best * 1e3, conv / best, conv, r["iters"].mean(),
    r["fevals"].sum() / r["iters"].sum(), h.hexdigest()[:16]))

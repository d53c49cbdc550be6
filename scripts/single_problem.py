"""BASELINE.json configs[0]: ONE Rosenbrock n = 60 problem from a randn start (README.md:52-59 of the
reference) through the host entry qn_bfgs_optimize_f64 — median wall time of repeated solves, beside
the CPU ports on this box's host and the reference's published 330.8 µs (README.md:80)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qn_b200 as qn
from oracle import binding as ob
ob.build()
n = 60
rng = np.random.Generator(np.random.Philox(key=20261018))
x0 = rng.standard_normal(n)
st = qn.BFGSState(n)
for _ in range(20):
    qn.optimize_(st, qn.Rosenbrock(), x0)
lat = []
for _ in range(300):
    t = time.perf_counter(); ell = qn.optimize_(st, qn.Rosenbrock(), x0); lat.append(time.perf_counter() - t)
ref = ob.optimize(ob.rosenbrock(), x0)
assert ell == ref["ell"] and np.array_equal(qn.optimum(st), ref["x"])
fast = ob.fast_time_single(x0, reps=300)
strict = []
for _ in range(100):
    t = time.perf_counter(); ob.optimize(ob.rosenbrock(), x0); strict.append(time.perf_counter() - t)
print("single problem n=60, %d iterations: GPU host entry median %.1f us (min %.1f) | CPU fast port %.1f us | CPU strict oracle %.1f us | reference README.md:80: 330.8 us (unstated CPU)" % (
    st.iters, np.median(lat) * 1e6, min(lat) * 1e6, np.median(fast) * 1e6, np.median(strict) * 1e6))

"""First-light check on a B200: bitwise parity vs the oracle + rough timing."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qn_b200 as qn
from oracle import binding as ob

eng = qn.default_engine()
print(eng.device_info())
rng = np.random.Generator(np.random.Philox(key=20261018))
res = {}

def compare(n, batch, order, flags, ntrace=5):
    Theta = rng.standard_normal((batch, n)).T.copy(order="F")
    ls = qn.BackTracking(order=order)
    r = eng.optimize_batched(qn.Rosenbrock(), Theta, ls=ls, ntrace=ntrace, flags=flags)
    o = ob.optimize_batch(ob.rosenbrock(), Theta, ls=ob.backtracking(order=order), ntrace=ntrace)
    same_x = np.array_equal(r["x"], o["x"])
    same_g = np.array_equal(r["g"], o["g"])
    same_it = np.array_equal(r["iters"], o["iters"]) and np.array_equal(r["fevals"], o["fevals"])
    same_st = np.array_equal(r["status"], o["status"])
    same_ell = np.array_equal(r["ell"], o["ell"], equal_nan=True)
    same_tr = np.array_equal(r["trace"], o["trace"], equal_nan=True)
    dx = np.nanmax(np.abs(r["x"] - o["x"]))
    print("n=%d batch=%d order=%d flags=%d path=%s: x %s g %s iters %s status %s ell %s trace %s  max|dx|=%.3e  conv=%d/%d iters mean %.1f" % (
        n, batch, order, flags, eng.last_launch(), same_x, same_g, same_it, same_st, same_ell, same_tr, dx,
        (r["status"] == 0).sum(), batch, r["iters"].mean()))
    if not same_it:
        bad = np.nonzero(r["iters"] != o["iters"])[0][:5]
        print("   first mismatches", bad, r["iters"][bad], o["iters"][bad])
        if ntrace:
            j = bad[0]
            print("   trace diff per iter:", np.nanmax(np.abs(r["trace"][j] - o["trace"][j]), axis=1))
    return bool(same_x and same_g and same_it and same_st and same_ell and same_tr)

ok = True
ok &= compare(60, 256, 2, 0)
ok &= compare(60, 256, 2, qn._lib.FLAG_FORCE_SMEM)
ok &= compare(60, 256, 3, 0)
for n in (2, 3, 5, 8, 16, 17, 24, 32, 33, 61, 64, 100, 128):
    ok &= compare(n, 64, 2, 0)
res["parity_ok"] = bool(ok)
print("PARITY", ok)

# timing
import ctypes as C
for flags, name in ((0, "reg"), (qn._lib.FLAG_FORCE_SMEM, "smem")):
    n, batch = 60, 65536
    Theta = rng.standard_normal((batch, n)).T.copy(order="F")
    for rep in range(3):
        t = time.time()
        r = eng.optimize_batched(qn.Rosenbrock(), Theta, flags=flags)
        dt = time.time() - t
    its = int(r["iters"].sum())
    print("%s: e2e %.1f ms  -> %.0f solves/s  %.3e iters/s  conv %d  launch %s" % (
        name, dt * 1e3, batch / dt, its / dt, (r["status"] == 0).sum(), eng.last_launch()))
    res[name] = dict(ms=dt * 1e3, solves_per_s=batch / dt, iters_per_s=its / dt)
mb = eng.microbench()
print("microbench", mb)
res["microbench"] = mb
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/first_light.json", "w"), indent=1)

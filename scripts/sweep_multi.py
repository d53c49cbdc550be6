"""BASELINE.json configs[2] and [3] on G GPUs of one box, through the library's own multi-GPU entry.

    python scripts/sweep_multi.py --gpus G [--quick] [--no-cpu]     # writes gpurun_out/r2_sweep_{G}gpu.json

Config 3 (`configs[2]`): dense-Gaussian and logistic-regression log-densities, n = 128, 2^15 problems.
Config 4 (`configs[3]`): Rosenbrock, n in {8, 16, 32, 60, 128, 160} x batch in {2^10 .. 2^18}.
One process drives all G GPUs with qn_bfgs_optimize_batched_multi_f64 (problem j -> GPU j mod G, no
inter-GPU traffic): pinned HOST buffers, so every number is end to end — host scatter, solve, host
gather — by the wall clock around the call (it returns synchronised).  Beside each row: the CPU arm
on the box's host cores (fast port for Rosenbrock, strict oracle for the build-defined objectives; a
bounded sample) and the FP64 roofline fraction (8 n² flops per iteration x iterations / time against
G x the DFMA rate measured live).  Rows whose estimated time exceeds --row-seconds are skipped.
"""
import argparse, json, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import qn_b200 as qn
from problems import gaussian_problem, logistic_problem

ap = argparse.ArgumentParser()
ap.add_argument("--gpus", type=int, default=1)
ap.add_argument("--quick", action="store_true")
ap.add_argument("--no-cpu", action="store_true")
ap.add_argument("--row-seconds", type=float, default=4.0)
ap.add_argument("--objectives-only", action="store_true", help="only the n = 128 dense-Gaussian / logistic rows; merged into the existing JSON")
args = ap.parse_args()
G = args.gpus
assert qn._lib.load().qn_device_count() >= G, "not enough GPUs"
engines = [qn.Engine(g) for g in range(G)]
dfma = engines[0].microbench()["dfma"]  # TFLOP/s per GPU, measured live
print("G=%d  DFMA peak per GPU %.2f TFLOP/s" % (G, dfma), flush=True)
if not args.no_cpu:
    from oracle import binding as ob
    ob.build()
    cores = ob.num_cores()


def pinned(shape, dtype):
    return torch.empty(shape, dtype=dtype).pin_memory()


def run(obj, n, batch, reps=2):
    rng = np.random.Generator(np.random.Philox(key=20261018 + n))
    th = pinned((batch, n), torch.float64)
    th.numpy()[...] = rng.standard_normal((batch, n))
    Theta = th.numpy().T  # n x batch, column-major view of the pinned buffer
    out = dict(x=pinned((batch, n), torch.float64).numpy().T, g=pinned((batch, n), torch.float64).numpy().T,
               ell=pinned((batch,), torch.float64).numpy(), iters=pinned((batch,), torch.int32).numpy(),
               fevals=pinned((batch,), torch.int32).numpy(), status=pinned((batch,), torch.int32).numpy())
    qn.optimize_batched_multi(engines, obj, Theta, out=out)  # warm-up (allocation, first launch)
    best = 1e30
    for _ in range(reps):
        t = time.perf_counter()
        r = qn.optimize_batched_multi(engines, obj, Theta, out=out)
        best = min(best, time.perf_counter() - t)
    conv = int((r["status"] == 0).sum()); iters = int(r["iters"].sum())
    its = iters / best
    return dict(n=n, batch=batch, gpus=G, ms=best * 1e3, solves_per_s=conv / best, iters_per_s=its, converged=conv,
                mean_iters=iters / batch, path=engines[0].last_launch()["path"],
                fp64_roofline_frac=its * 8.0 * n * n / 1e12 / (dfma * G)), Theta


def cpu_rosenbrock(n, Theta):
    """fast port, all host cores, a sample sized for about 0.4 s"""
    s = min(Theta.shape[1], 2 * cores)
    t = time.perf_counter(); ob.fast_optimize_batch(np.asfortranarray(Theta[:, :s]), want_xg=False); dt = time.perf_counter() - t
    s = int(min(Theta.shape[1], max(s, s * 0.4 / max(dt, 1e-4))))
    t = time.perf_counter(); r = ob.fast_optimize_batch(np.asfortranarray(Theta[:, :s]), want_xg=False); dt = time.perf_counter() - t
    return dict(cpu_solves_per_s=r["converged"] / dt, cpu_sample=s, cpu_cores=cores, cpu_kind="fast port (oracle/qn_fastport.c)")


rows = {"config": "BASELINE.json configs[2], configs[3]", "gpus": G, "dfma_tflops_per_gpu": dfma,
        "api": "qn_bfgs_optimize_batched_multi_f64, pinned host buffers, wall clock around the call",
        "rosenbrock": [], "objectives": [], "skipped": []}
# measured round-1 single-GPU rates (iterations/s) to predict a row's time and skip the hopeless ones
rate1 = {8: 8.9e9, 16: 4.0e9, 32: 1.1e9, 60: 3.9e8, 128: 7.0e7, 160: 2.2e7}
iters1 = {8: 103, 16: 172, 32: 280, 60: 435, 128: 723, 160: 835}
ns = [] if args.objectives_only else [8, 16, 32, 60, 128, 160]
cpu_done = set()
batches = [1 << 10, 1 << 14] if args.quick else [1 << k for k in range(10, 19)]
for n in ns:
    for b in batches:
        est = b * iters1[n] / (rate1[n] * G)
        if est > args.row_seconds:
            rows["skipped"].append(dict(n=n, batch=b, estimated_s=est))
            continue
        r, Theta = run(qn.Rosenbrock(), n, b)
        if not args.no_cpu and n not in cpu_done:  # per-solve CPU cost does not depend on the batch: once per n
            cpu_done.add(n)
            r.update(cpu_rosenbrock(n, Theta))
        rows["rosenbrock"].append(r)
        print("rosenbrock n=%3d batch=%7d G=%d: %9.2f ms  %11.0f solves/s  %.3e it/s  conv %d  frac %.3f %s" % (
            n, b, G, r["ms"], r["solves_per_s"], r["iters_per_s"], r["converged"], r["fp64_roofline_frac"],
            ("cpu %.0f solves/s on %d cores" % (r["cpu_solves_per_s"], r["cpu_cores"])) if "cpu_solves_per_s" in r else ""), flush=True)
# one CPU figure per n is enough (per-solve cost does not depend on the batch): fill it in for every row of that n
cpu_by_n = {r["n"]: {k: r[k] for k in r if k.startswith("cpu_")} for r in rows["rosenbrock"] if "cpu_solves_per_s" in r}
for r in rows["rosenbrock"]:
    r.update(cpu_by_n.get(r["n"], {}))
    if "cpu_solves_per_s" in r:
        r["speedup_vs_cpu_all_cores"] = r["solves_per_s"] / r["cpu_solves_per_s"]

mu, lam = gaussian_problem(128)
X, y, l2 = logistic_problem(1024, 128)
for name, obj in (("dense_gaussian", qn.DenseGaussian(mu, lam)), ("logistic", qn.LogisticRegression(X, y, l2))):
    for b in ([1 << 10] if args.quick else [1 << 12, 1 << 15]):
        r, Theta = run(obj, 128, b)
        r["objective"] = name
        r["fp64_roofline_frac_note"] = "BFGS part only (8 n^2 flops per iteration); the objective's own GEMVs are not counted"
        if not args.no_cpu:
            om = ob.OracleModel(ob.DENSE_GAUSSIAN, mu=mu, lam=lam) if name == "dense_gaussian" else ob.OracleModel(ob.LOGISTIC, X=X, y=y, lam=l2)
            sample = np.asfortranarray(Theta[:, :16 * cores])
            t = time.perf_counter(); o = ob.optimize_batch(om, sample); dt = time.perf_counter() - t
            r.update(cpu_solves_per_s=o["converged"] / dt, cpu_sample=sample.shape[1], cpu_cores=cores,
                     cpu_kind="strict oracle (oracle/qn_oracle.c)")
            r["speedup_vs_cpu_all_cores"] = r["solves_per_s"] / r["cpu_solves_per_s"]
        rows["objectives"].append(r)
        print("%s n=128 batch=%d G=%d: %.2f ms  %.0f solves/s  conv %d/%d  iters/solve %.1f %s" % (
            name, b, G, r["ms"], r["solves_per_s"], r["converged"], b, r["mean_iters"],
            ("cpu %.0f solves/s on %d cores" % (r["cpu_solves_per_s"], r["cpu_cores"])) if "cpu_solves_per_s" in r else ""), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
if args.objectives_only:  # keep the Rosenbrock rows of the committed sweep, replace the objectives
    prev = os.path.join(ROOT, "profiles", "r2_sweep_%dgpu.json" % G)
    if os.path.exists(prev):
        old = json.load(open(prev))
        rows["rosenbrock"], rows["skipped"] = old.get("rosenbrock", []), old.get("skipped", [])
        rows["note"] = "objectives re-measured after the logistic GEMV rewrite; Rosenbrock rows from the earlier run of this round"
json.dump(rows, open("gpurun_out/r2_sweep_%dgpu.json" % G, "w"), indent=1)
print("wrote gpurun_out/r2_sweep_%dgpu.json" % G)

"""BASELINE.json configs[3]: multi-start sweep n x batch (Rosenbrock) + configs[2] objectives at n=128.

    python scripts/sweep.py [--quick] [--cpu]      # writes gpurun_out/sweep.json

Device-resident inputs, CUDA-event timing through the device-pointer C entry (torch holds the
buffers).  --cpu also times the CPU port (oracle, all host cores) on a bounded sample.
"""
import argparse, json, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import qn_b200 as qn
from problems import gaussian_problem, logistic_problem

ap = argparse.ArgumentParser()
ap.add_argument("--quick", action="store_true")
ap.add_argument("--cpu", action="store_true")
args = ap.parse_args()
eng = qn.Engine(0)
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream()


def run(obj, n, batch, reps=3):
    rng = np.random.Generator(np.random.Philox(key=20261018 + n))
    th = torch.from_numpy(rng.standard_normal((batch, n))).to(dev)
    x = torch.empty_like(th); g = torch.empty_like(th)
    ell = torch.empty(batch, dtype=torch.float64, device=dev)
    it = torch.empty(batch, dtype=torch.int32, device=dev); fe = torch.empty_like(it); st = torch.empty_like(it)
    def step():
        eng.optimize_batched_device(obj, n, batch, th.data_ptr(), x_ptr=x.data_ptr(), g_ptr=g.data_ptr(), ell_ptr=ell.data_ptr(),
                                    iters_ptr=it.data_ptr(), fevals_ptr=fe.data_ptr(), status_ptr=st.data_ptr(), stream=stream.cuda_stream)
    step(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e30
    for _ in range(reps):
        e0.record(stream); step(); e1.record(stream); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    conv = int((st == 0).sum()); iters = int(it.sum())
    return dict(n=n, batch=batch, ms=best, solves_per_s=conv / (best * 1e-3), iters_per_s=iters / (best * 1e-3),
                converged=conv, mean_iters=iters / batch, path=eng.last_launch()["path"], grid=eng.last_launch()["grid"],
                host_copy=th.cpu().numpy())

out = {"rosenbrock": [], "objectives": []}
ns = [8, 16, 32, 60, 128, 160]
batches = [1 << 10, 1 << 14] if args.quick else [1 << k for k in range(10, 19, 2)]
for n in ns:
    for b in batches:
        if n >= 128 and b > (1 << 14):
            continue  # hours of iterations at 1 CTA/SM; the trend is flat beyond one wave
        r = run(qn.Rosenbrock(), n, b)
        r.pop("host_copy")
        out["rosenbrock"].append(r)
        print("rosenbrock n=%3d batch=%7d: %9.2f ms  %10.0f solves/s  %.3e it/s  conv %d  iters/solve %.0f path %d" % (
            n, b, r["ms"], r["solves_per_s"], r["iters_per_s"], r["converged"], r["mean_iters"], r["path"]), flush=True)
mu, lam = gaussian_problem(128)
X, y, l2 = logistic_problem(1024, 128)
for name, obj in (("dense_gaussian", qn.DenseGaussian(mu, lam)), ("logistic", qn.LogisticRegression(X, y, l2))):
    for b in ([1 << 10] if args.quick else [1 << 10, 1 << 15]):
        r = run(obj, 128, b)
        hc = r.pop("host_copy")
        r["objective"] = name
        if args.cpu:
            from oracle import binding as ob
            om = ob.OracleModel(ob.DENSE_GAUSSIAN, mu=mu, lam=lam) if name == "dense_gaussian" else ob.OracleModel(ob.LOGISTIC, X=X, y=y, lam=l2)
            sample = hc[:256].T.copy(order="F")
            t = time.perf_counter(); o = ob.optimize_batch(om, sample); dt = time.perf_counter() - t
            r["cpu_port_solves_per_s"] = o["converged"] / dt; r["cpu_cores"] = ob.num_cores()
        out["objectives"].append(r)
        print("%s n=128 batch=%d: %.2f ms  %.0f solves/s  conv %d/%d  iters/solve %.1f %s" % (
            name, b, r["ms"], r["solves_per_s"], r["converged"], b, r["mean_iters"],
            ("cpu port %.0f solves/s on %d cores" % (r["cpu_port_solves_per_s"], r["cpu_cores"])) if args.cpu else ""), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/sweep.json", "w"), indent=1)

"""Row-sharded large-n mode: ms per BFGS iteration of ONE n-dimensional Rosenbrock problem on 1..G GPUs
(BASELINE.json configs[4]; qn_bfgs_optimize_rowsharded_f64), both sweep orders, with a parity check of
the sharded result against the single-GPU result on a short run.

    python scripts/rowsharded_bench.py [n] [iters]
"""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qn_b200 as qn

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 42
ndev = qn._lib.load().qn_device_count()
rng = np.random.Generator(np.random.Philox(key=20261018 + n))
theta = rng.standard_normal(n)
out = []
base = {}
for G in [g for g in (1, 2, 4, 8) if g <= ndev]:
    engs = [qn.Engine(g) for g in range(G)]
    for mode, flags in (("two-pass", 0), ("one-pass", qn._lib.FLAG_ONE_PASS)):
        chk = qn.optimize_rowsharded(engs, qn.Rosenbrock(), theta, max_iter=6, ntrace=6, flags=flags)
        if G == 1:
            base[mode] = chk
        same = all(np.array_equal(chk[k], base[mode][k], equal_nan=True) for k in ("x", "g", "trace", "fevals"))
        r = qn.optimize_rowsharded(engs, qn.Rosenbrock(), theta, max_iter=iters, flags=flags, bench_iters=iters)
        ms = r["ms_per_iter"]
        gbs = (16.0 * n * n + 96.0 * n) / (ms * 1e-3) / 1e9
        rec = dict(n=n, gpus=G, mode=mode, ms_per_iter=ms, algorithmic_gbs_total=gbs, same_bits_as_1gpu=bool(same))
        out.append(rec)
        print(rec, flush=True)
    for e in engs:
        e.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/rowsharded_n%d.json" % n, "w"), indent=1)

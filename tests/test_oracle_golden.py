"""The oracle against the committed fixtures (tests/golden/, made by make_golden.py).

rosenbrock_mp_trace.json is an independent 60-digit mpmath restatement of the reference
algorithm: the oracle's first five iterates must agree with it to 1e-12 relative — the
north-star gate — which shows the canonical summation order chosen in the oracle does
not matter at that gate.  oracle_vectors.npz pins the oracle's own bits.
"""
import json
import os

import numpy as np

from conftest import GOLDEN

REL = 1e-12  # BASELINE.json: "the first 5 iterates agree to 1e-12 relative"


def test_first_five_iterates_match_mpmath(ob):
    d = json.load(open(os.path.join(GOLDEN, "rosenbrock_mp_trace.json")))
    assert len(d["cases"]) >= 10
    for case in d["cases"]:
        x0 = np.array([float.fromhex(h) for h in case["x0"]])
        r = ob.optimize(ob.rosenbrock(), x0, ls=ob.backtracking(order=case["order"]), ntrace=5)
        ref = np.array(case["trace"])
        scale = np.maximum(np.abs(ref).max(axis=1, keepdims=True), 1e-300)
        rel = np.abs(r["trace"] - ref) / scale
        assert rel.max() <= REL, (case["n"], case["order"], rel.max())
        assert np.allclose(r["trace_scalars"][:, 2], case["alpha"], rtol=1e-9)


def test_oracle_bits_are_pinned(ob):
    v = np.load(os.path.join(GOLDEN, "oracle_vectors.npz"))
    keys = sorted({k.rsplit("_", 1)[0] for k in v.files})
    assert len(keys) == 16
    for k in keys:
        n, order = int(k.split("_")[0][1:]), int(k.split("_")[1][1:])
        r = ob.optimize_batch(ob.rosenbrock(), v[k + "_theta0"], ls=ob.backtracking(order=order), ntrace=5)
        for f in ("x", "g", "ell", "iters", "fevals", "status", "trace"):
            assert np.array_equal(r[f], v[k + "_" + f], equal_nan=True), (k, f)

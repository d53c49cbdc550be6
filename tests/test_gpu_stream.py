"""The large-n path (BASELINE.json configs[4]): B⁻¹ in HBM, fused streaming GEMV / rank-2 passes.
Same canonical order as the resident kernels, so the comparison with the oracle is bit for bit."""
import numpy as np
import pytest

from conftest import starts

pytestmark = pytest.mark.gpu


def _same(got, ref):
    for k in ("x", "g", "ell", "trace"):
        if ref.get(k) is not None:
            assert np.array_equal(got[k], ref[k], equal_nan=True), k
    assert np.array_equal(got["iters"], ref["iters"])
    assert np.array_equal(got["fevals"], ref["fevals"])
    assert np.array_equal(got["status"], ref["status"])


@pytest.mark.parametrize("n", [2, 17, 60, 61])
def test_forced_stream_path_full_solves(eng, qn, ob, n):
    """QN_FLAG_FORCE_STREAM on sizes the resident path also covers: identical bits, full solves."""
    Theta = starts(n, 3, key=900 + n)
    got = eng.optimize_batched(qn.Rosenbrock(), Theta, ntrace=5, flags=qn._lib.FLAG_FORCE_STREAM)
    assert eng.last_launch()["path"] == qn._lib.PATH_STREAM
    ref = ob.optimize_batch(ob.rosenbrock(), Theta, ntrace=5)
    _same(got, ref)
    assert (got["status"] == qn.CONVERGED).all()


@pytest.mark.parametrize("n,iters", [(161, 40), (256, 40), (777, 25), (1024, 25), (2049, 12)])
def test_large_n_first_iterations(eng, qn, ob, n, iters):
    """n beyond the resident limit: the first `iters` iterations, iterates and all, against the
    oracle (a full solve at these sizes takes thousands of iterations; QN.jl:250 caps at 10 000)."""
    Theta = starts(n, 1, key=n)
    got = eng.optimize_batched(qn.Rosenbrock(), Theta, max_iter=iters, ntrace=iters)
    assert eng.last_launch()["path"] == qn._lib.PATH_STREAM
    ref = ob.optimize_batch(ob.rosenbrock(), Theta, max_iter=iters, ntrace=iters)
    _same(got, ref)
    assert got["status"][0] == qn.MAX_ITER and np.isnan(got["ell"][0])


def test_large_n_full_solve_and_h_writeback(eng, qn, ob):
    n = 192
    Theta = starts(n, 1, key=5)
    sb = qn.BFGSStateBatch(n, 1)
    got = eng.optimize_batched(qn.Rosenbrock(), Theta, states=sb, flags=qn._lib.FLAG_WRITE_H)
    o = ob.optimize(ob.rosenbrock(), Theta[:, 0], L=sb.L)
    assert got["status"][0] == qn.CONVERGED and got["iters"][0] == o["iters"]
    assert np.array_equal(got["x"][:, 0], o["x"]) and got["ell"][0] == o["ell"]
    L = sb.L
    assert np.array_equal(sb.data[0, 7 * L:].reshape(n, L)[:, :n], o["state"][7 * L:].reshape(n, L)[:, :n])
    print("n=192 full solve on the streaming path: %d iterations" % got["iters"][0])


def test_stream_rejects_unregistered_objective(eng, qn):
    """An objective without a device implementation on this path is an error, never a CPU fallback
    (the MethodError analogue, QN.jl:254): logistic regression has none beyond the resident sizes."""
    from problems import logistic_problem
    X, y, lam = logistic_problem(64, 200)
    with pytest.raises(qn.UnregisteredObjectiveError):
        eng.optimize_batched(qn.LogisticRegression(X, y, lam), starts(200, 1))


@pytest.mark.parametrize("n,iters", [(60, 0), (200, 0), (333, 0), (700, 40)])
def test_stream_dense_gaussian(eng, qn, ob, n, iters):
    """The reference's optimize! takes any objective at any P (QN.jl:254); here the dense Gaussian runs
    beyond the resident sizes through a CTA-cooperative objective in the streaming path's vector
    kernels (rule-C2 row chains for Λ(θ-μ)): bit for bit its oracle, full solves to θ* = μ."""
    from problems import gaussian_problem
    mu, lam = gaussian_problem(n)
    model = qn.DenseGaussian(mu, lam)
    omodel = ob.OracleModel(ob.DENSE_GAUSSIAN, mu=mu, lam=lam)
    Theta = starts(n, 2, key=5200 + n)
    kw = dict(max_iter=iters) if iters else {}
    nt = iters if iters else 5
    flags = qn._lib.FLAG_FORCE_STREAM if n <= 160 else 0
    got = eng.optimize_batched(model, Theta, ntrace=nt, flags=flags, **kw)
    assert eng.last_launch()["path"] == qn._lib.PATH_STREAM
    ref = ob.optimize_batch(omodel, Theta, ntrace=nt, **kw)
    _same(got, ref)
    if not iters:
        assert (got["status"] == qn.CONVERGED).all()
        assert np.abs(got["x"] - mu[:, None]).max() < 1e-7
    # QN_FLAG_ONE_PASS is a permission: this objective keeps the reference order
    got1 = eng.optimize_batched(model, Theta[:, :1], ntrace=nt, flags=flags | qn._lib.FLAG_ONE_PASS, **kw)
    assert np.array_equal(got1["x"][:, 0], got["x"][:, 0]) and got1["iters"][0] == got["iters"][0]


@pytest.mark.parametrize("n,iters", [(60, 0), (61, 0), (200, 60), (1024, 20), (8193, 6), (9001, 6), (16500, 4)])
def test_one_pass_mode(eng, qn, ob, n, iters):
    """QN_FLAG_ONE_PASS (deferred update, 16 n² bytes per iteration): bit for bit equal to ITS oracle
    (oracle rblock = ONEPASS), and within the north-star tolerances of the reference order:
    first five iterates 1e-12 relative, max|∇| < 1e-8; the optimum agrees to 5e-8 (two valid runs
    stopped at max|∇| < 1e-8 can differ by ~2·tol/λ_min = 5e-8 on Rosenbrock)."""
    flags = qn._lib.FLAG_FORCE_STREAM | qn._lib.FLAG_ONE_PASS
    Theta = starts(n, 2, key=70 + n)
    kw = dict(max_iter=iters) if iters else {}
    nt = iters if iters else 5
    got = eng.optimize_batched(qn.Rosenbrock(), Theta, ntrace=nt, flags=flags, **kw)
    own = ob.optimize_batch(ob.rosenbrock(), Theta, ntrace=nt, rblock=ob.ONEPASS, **kw)
    _same(got, own)
    ref = ob.optimize_batch(ob.rosenbrock(), Theta, ntrace=nt, **kw)
    scale = np.abs(ref["trace"][:, :5]).max(axis=2, keepdims=True)
    assert (np.abs(got["trace"][:, :5] - ref["trace"][:, :5]) / scale).max() <= 1e-12
    if not iters:
        assert (got["status"] == qn.CONVERGED).all()
        assert np.abs(got["g"]).max() < 1e-8
        assert np.abs(got["x"] - ref["x"]).max() <= 5e-8
        print("one-pass n=%d: iterations %s vs reference order %s" % (n, got["iters"], ref["iters"]))


@pytest.mark.parametrize("n,iters", [(61, 0), (200, 60)])
def test_one_pass_mode_single_cta_fallback(eng, qn, ob, n, iters, option):
    """The single-CTA vector kernel (used when n is beyond the cluster kernel's reach) gives the same bits."""
    option("stream_cluster", 0)
    flags = qn._lib.FLAG_FORCE_STREAM | qn._lib.FLAG_ONE_PASS
    Theta = starts(n, 2, key=70 + n)
    kw = dict(max_iter=iters) if iters else {}
    nt = iters if iters else 5
    got = eng.optimize_batched(qn.Rosenbrock(), Theta, ntrace=nt, flags=flags, **kw)
    own = ob.optimize_batch(ob.rosenbrock(), Theta, ntrace=nt, rblock=ob.ONEPASS, **kw)
    _same(got, own)


def test_one_pass_h_writeback(eng, qn, ob):
    n = 40
    Theta = starts(n, 1, key=8)
    sb = qn.BFGSStateBatch(n, 1)
    eng.optimize_batched(qn.Rosenbrock(), Theta, states=sb,
                         flags=qn._lib.FLAG_WRITE_H | qn._lib.FLAG_FORCE_STREAM | qn._lib.FLAG_ONE_PASS)
    o = ob.optimize(ob.rosenbrock(), Theta[:, 0], L=sb.L, rblock=ob.ONEPASS)
    L = sb.L
    assert np.array_equal(sb.data[0, 7 * L:].reshape(n, L)[:, :n], o["state"][7 * L:].reshape(n, L)[:, :n])


@pytest.mark.parametrize("n,iters", [(17, 0), (61, 0), (192, 0), (161, 40), (777, 25), (2049, 12), (8192, 5)])
def test_tma_staged_sweeps_same_bits(eng, qn, ob, n, iters, option):
    """Option stream_tma = 1: the two sweeps of the reference-order mode staged through a
    shared-memory ring by the bulk-copy engine (cp.async.bulk + mbarrier; stream_pass1_tma /
    stream_pass2_tma).  Same arithmetic in the same order: bit for bit the oracle — full solves
    (with their resets of B⁻¹), ragged sizes, the final B⁻¹ the engine stored, and config 5's size."""
    option("stream_tma", 1)
    Theta = starts(n, 1, key=4400 + n)
    kw = dict(max_iter=iters) if iters else {}
    nt = iters if iters else 5
    sb = qn.BFGSStateBatch(n, 1) if n <= 800 else None
    flags = qn._lib.FLAG_FORCE_STREAM | (qn._lib.FLAG_WRITE_H if sb is not None else 0)
    got = eng.optimize_batched(qn.Rosenbrock(), Theta, ntrace=nt, flags=flags, states=sb, **kw)
    assert eng.last_launch()["path"] == qn._lib.PATH_STREAM
    ref = ob.optimize_batch(ob.rosenbrock(), Theta, ntrace=nt, **kw)
    _same(got, ref)
    if sb is not None:
        L = sb.L
        o = ob.optimize(ob.rosenbrock(), Theta[:, 0], L=L, **kw)
        assert np.array_equal(sb.data[0, 7 * L:].reshape(n, L)[:, :n], o["state"][7 * L:].reshape(n, L)[:, :n])

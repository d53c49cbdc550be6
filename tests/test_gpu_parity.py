"""Parity of the CUDA path (through the C ABI) with the CPU oracle.  -m gpu.

The north-star gates (BASELINE.json): converged optimum ‖Δθ‖∞ <= 1e-8, max|∇| < 1e-8,
first 5 iterates to 1e-12 relative, iteration counts side by side.  Because the kernels
follow the oracle's canonical arithmetic order the comparison here is
stricter: BIT-FOR-BIT equality of θ*, ∇, ℓ, the traces and the iteration/evaluation
counts.  The tolerance gates are asserted as well, so a future kernel that gives up
bit-exactness still has to meet them.
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN, starts

pytestmark = pytest.mark.gpu

TOL_X = 1e-8       # ‖Δθ‖∞ at the optimum
TOL_G = 1e-8       # max|∇|
TOL_TRACE = 1e-12  # first five iterates, relative


def _check(got, ref, bitwise=True):
    assert np.array_equal(got["status"], ref["status"])
    conv = ref["status"] == 0
    # the gates north_star names
    assert np.nanmax(np.abs(got["x"][:, conv] - ref["x"][:, conv]), initial=0) <= TOL_X
    assert np.abs(got["g"][:, conv]).max(initial=0) < TOL_G
    if ref.get("trace") is not None:
        scale = np.maximum(np.nanmax(np.abs(ref["trace"]), axis=2, keepdims=True), 1e-300)
        rel = np.abs(got["trace"] - ref["trace"]) / scale
        assert np.nanmax(rel, initial=0) <= TOL_TRACE
        assert np.array_equal(np.isnan(got["trace"]), np.isnan(ref["trace"]))
    if bitwise:
        for k in ("x", "g", "ell"):
            assert np.array_equal(got[k], ref[k], equal_nan=True), k
        if ref.get("trace") is not None:
            assert np.array_equal(got["trace"], ref["trace"], equal_nan=True)
        assert np.array_equal(got["iters"], ref["iters"])
        assert np.array_equal(got["fevals"], ref["fevals"])


@pytest.mark.parametrize("n", list(range(2, 25)))
@pytest.mark.parametrize("order", [2, 3])
def test_reference_test_range(eng, qn, ob, n, order):
    """The sizes of the reference's own testset (test/runtests.jl:37): n = 2..24, both orders."""
    Theta = starts(n, 48, key=100 * order + n)
    got = eng.optimize_batched(qn.Rosenbrock(), Theta, ls=qn.BackTracking(order=order), ntrace=5)
    ref = ob.optimize_batch(ob.rosenbrock(), Theta, ls=ob.backtracking(order=order), ntrace=5)
    _check(got, ref)
    assert (got["status"] == qn.CONVERGED).all()
    assert np.abs(got["x"] - 1).max() < 1e-7


@pytest.mark.parametrize("order", [2, 3])
@pytest.mark.parametrize("path", ["register", "smem"])
def test_n60_both_resident_paths(eng, qn, ob, order, path):
    """README.md:52-75 size, batched; the register- and smem-resident kernels give the same bits."""
    Theta = starts(60, 1024, key=60 + order)
    flags = qn._lib.FLAG_FORCE_SMEM if path == "smem" else 0
    got = eng.optimize_batched(qn.Rosenbrock(), Theta, ls=qn.BackTracking(order=order), ntrace=5, flags=flags)
    assert eng.last_launch()["path"] == (qn._lib.PATH_SMEM if path == "smem" else qn._lib.PATH_REGISTER)
    ref = ob.optimize_batch(ob.rosenbrock(), Theta, ls=ob.backtracking(order=order), ntrace=5)
    _check(got, ref)
    assert (got["status"] == qn.CONVERGED).all()
    print("n=60 order=%d %s: iterations GPU mean %.2f / oracle mean %.2f, max %d / %d" % (
        order, path, got["iters"].mean(), ref["iters"].mean(), got["iters"].max(), ref["iters"].max()))


@pytest.mark.parametrize("n", [1, 31, 32, 33, 59, 61, 63, 64, 65, 96, 100, 127, 128, 129, 160])
def test_ragged_sizes_smem_path(eng, qn, ob, n):
    """Odd n (tail term), warp-boundary sizes, and the largest resident size (n=160, 205 KB of B⁻¹)."""
    batch = 12 if n > 100 else 24
    Theta = starts(n, batch, key=n)
    got = eng.optimize_batched(qn.Rosenbrock(), Theta, ntrace=5)
    ref = ob.optimize_batch(ob.rosenbrock(), Theta, ntrace=5)
    _check(got, ref)


@pytest.mark.parametrize("n,variant", [(8, "1"), (8, "2"), (8, "4"), (16, "1"), (16, "2"), (16, "4"), (32, "2")])
@pytest.mark.parametrize("order", [2, 3])
def test_packed_small_n(eng, qn, ob, n, variant, order, option):
    """n = 8, 16: several problems per warp (qn_resident_packed.cuh), ragged batch sizes so that
    sub-groups run dry at different times, iteration caps, traces; and the thread-per-row kernel
    (option pack = 0) gives the same bits."""
    option("pack_variant", int(variant))  # rows per lane (qn_capi.cu dispatch_size)
    ls = qn.BackTracking[order]()
    ols = ob.backtracking(order)
    for batch, kw in ((1, {}), (7, {}), (133, {}), (64, dict(max_iter=3)), (64, dict(max_iter=4))):
        Theta = starts(n, batch, key=1000 * n + batch)
        ref = ob.optimize_batch(ob.rosenbrock(), Theta, ls=ols, ntrace=5, **kw)
        got = eng.optimize_batched(qn.Rosenbrock(), Theta, ls=ls, ntrace=5, **kw)
        _check(got, ref)
        option("pack", 0)
        got2 = eng.optimize_batched(qn.Rosenbrock(), Theta, ls=ls, ntrace=5, **kw)
        option("pack", 1)
        _check(got2, ref)


@pytest.mark.parametrize("n", [36, 40, 44, 48, 50, 52, 56, 60, 64])
def test_tile_kernel_sizes(eng, qn, ob, n):
    """Every size with a 4×CT register-tile instantiation (qn_reg_sizes.inc), both line-search orders."""
    Theta = starts(n, 20, key=7000 + n)
    for order in (2, 3):
        got = eng.optimize_batched(qn.Rosenbrock(), Theta, ls=qn.BackTracking[order](), ntrace=5)
        assert eng.last_launch()["path"] == qn._lib.PATH_REGISTER and eng.last_launch()["block"] == 64
        ref = ob.optimize_batch(ob.rosenbrock(), Theta, ls=ob.backtracking(order), ntrace=5)
        _check(got, ref)


@pytest.mark.parametrize("n", [36, 40, 44, 48, 50, 52, 56, 60, 64])
def test_tile_kernel_in_tensor_memory(eng, qn, ob, n, option):
    """The same kernel with B⁻¹ tiles in tensor memory (tcgen05.ld/st; CTAs of 128 threads = two
    problems, 8 problems per SM): the default for large batches of n = 44, 48, 50, 56, 60, 64, forced
    here for every tile size (option tile_tmem = 1) and for batches that leave a pair of warps idle.  Same
    bits as the oracle — traces, counts, both line-search orders, the final B⁻¹, iteration caps
    (which ∇ the canonical slot shows), a non-finite start — and as the all-register kernel."""
    option("tile_tmem", 1)
    R = qn.Rosenbrock()
    for batch in (1, 21):
        Theta = starts(n, batch, key=8000 + n)
        for order in (2, 3):
            got = eng.optimize_batched(R, Theta, ls=qn.BackTracking[order](), ntrace=5)
            assert eng.last_launch()["path"] == qn._lib.PATH_REGISTER and eng.last_launch()["block"] in (128, 384)
            ref = ob.optimize_batch(ob.rosenbrock(), Theta, ls=ob.backtracking(order), ntrace=5)
            _check(got, ref)
    # final B⁻¹ (read back from tensor memory) and the BFGSState layout
    batch = 5
    Theta = starts(n, batch, key=8100 + n)
    sb = qn.BFGSStateBatch(n, batch)
    sb.data[:] = np.nan
    r = eng.optimize_batched(R, Theta, states=sb, flags=qn._lib.FLAG_WRITE_H)
    assert eng.last_launch()["block"] in (128, 384)
    L = sb.L
    for j in range(batch):
        o = ob.optimize(ob.rosenbrock(), Theta[:, j], L=L)
        assert np.array_equal(sb.data[j, 7 * L:].reshape(n, L)[:, :n], o["state"][7 * L:].reshape(n, L)[:, :n])
        assert np.array_equal(sb.data[j, :n], o["x"]) and np.array_equal(sb.data[j, L:L + n], o["g"])
    # iteration caps (odd: latest ∇, even: previous ∇ — kept in registers on this path) and a NaN start
    for cap in (1, 2, 5, 6):
        Th = starts(n, 4, key=cap)
        if cap == 5:
            Th[3, 2] = np.nan
        got = eng.optimize_batched(R, Th, max_iter=cap, ntrace=cap)
        ref = ob.optimize_batch(ob.rosenbrock(), Th, max_iter=cap, ntrace=cap)
        for k in ("x", "g", "trace", "ell"):
            assert np.array_equal(got[k], ref[k], equal_nan=True), (n, cap, k)
        for k in ("iters", "fevals", "status"):
            assert np.array_equal(got[k], ref[k]), (n, cap, k)
    # and the all-register kernel gives the same bits on a batch large enough for either
    Theta = starts(n, 300, key=8200 + n)
    a = eng.optimize_batched(R, Theta)
    option("tile_tmem", 0)
    b_ = eng.optimize_batched(R, Theta)
    assert eng.last_launch()["block"] == 64
    for k in ("x", "g", "ell", "iters", "fevals", "status"):
        assert np.array_equal(a[k], b_[k], equal_nan=True), k


ALL_PATH_SIZES = [1, 2, 3, 5, 7, 8, 9, 15, 16, 17, 31, 32, 33, 36, 40, 47, 48, 49, 56, 60, 63, 64, 65, 66,
                  80, 95, 96, 104, 105, 112, 127, 128, 129, 144, 159, 160]


@pytest.mark.parametrize("flags_name", ["default", "smem"])
def test_every_dispatch_branch_first_iterations(eng, qn, ob, flags_name):
    """One sweep over sizes chosen to hit every branch of the resident dispatch (packed, thread-per-row,
    register tiles of both kinds, shared memory with 1, 2 or 4 rows per thread, odd sizes with their
    padding rows): 30 iterations with every iterate traced, bit for bit against the oracle, for both
    line-search orders; and once more with QN_FLAG_FORCE_SMEM."""
    flags = qn._lib.FLAG_FORCE_SMEM if flags_name == "smem" else 0
    for n in ALL_PATH_SIZES:
        Theta = starts(n, 4, key=9000 + n)
        for order in (2, 3):
            got = eng.optimize_batched(qn.Rosenbrock(), Theta, ls=qn.BackTracking[order](), max_iter=30, ntrace=30, flags=flags)
            ref = ob.optimize_batch(ob.rosenbrock(), Theta, ls=ob.backtracking(order), max_iter=30, ntrace=30)
            for k in ("x", "g", "ell", "trace"):
                assert np.array_equal(got[k], ref[k], equal_nan=True), (n, order, k)
            for k in ("iters", "fevals", "status"):
                assert np.array_equal(got[k], ref[k]), (n, order, k)


def test_n128_both_resident_policies(eng, qn, ob):
    """n = 128 runs with B⁻¹ as 4x16 register tiles of 256 threads by default and in shared memory
    (row pairs, column blocks dealt out to thread groups) under QN_FLAG_FORCE_SMEM: same bits."""
    Theta = starts(128, 8, key=1280)
    ref = ob.optimize_batch(ob.rosenbrock(), Theta, ntrace=5)
    for flags, path in ((0, qn._lib.PATH_REGISTER), (qn._lib.FLAG_FORCE_SMEM, qn._lib.PATH_SMEM)):
        got = eng.optimize_batched(qn.Rosenbrock(), Theta, ntrace=5, flags=flags)
        assert eng.last_launch()["path"] == path
        _check(got, ref)


def test_n128_tiles_with_rows_in_tensor_memory(eng, qn, ob, option):
    """n = 128 hybrid tiles whose two non-register rows live in tensor memory instead of shared memory
    (option pair_rows = 't', the default): same bits as the oracle, including the final B⁻¹ read back from
    tensor memory, and as the shared-memory hybrid (pair_rows = 'h')."""
    n = 128
    Theta = starts(n, 40, key=1289)
    a = eng.optimize_batched(qn.Rosenbrock(), Theta)
    assert eng.last_launch()["smem_bytes"] < 32768  # no B⁻¹ rows in shared memory
    option("pair_rows", "h")
    b_ = eng.optimize_batched(qn.Rosenbrock(), Theta)
    assert eng.last_launch()["smem_bytes"] > 65536
    for k in ("x", "g", "ell", "iters", "fevals", "status"):
        assert np.array_equal(a[k], b_[k], equal_nan=True), k
    option("pair_rows", "t")
    for batch in (1, 5, 300):
        Theta = starts(n, batch, key=1290 + batch)
        nref = min(batch, 6)
        sb = qn.BFGSStateBatch(n, batch)
        got = eng.optimize_batched(qn.Rosenbrock(), Theta, ntrace=5, states=sb, flags=qn._lib.FLAG_WRITE_H)
        assert eng.last_launch()["path"] == qn._lib.PATH_REGISTER and eng.last_launch()["block"] == 256
        ref = ob.optimize_batch(ob.rosenbrock(), Theta[:, :nref], ntrace=5)
        sub = {k: (None if got[k] is None else got[k][:nref] if k == "trace" else got[k][..., :nref]) for k in got}
        _check(sub, ref)
        L = sb.L
        o = ob.optimize(ob.rosenbrock(), Theta[:, 0], L=L)
        assert np.array_equal(sb.data[0, 7 * L:].reshape(n, L)[:, :n], o["state"][7 * L:].reshape(n, L)[:, :n])
    for cap in (1, 2, 3):
        Th = starts(n, 3, key=cap)
        got = eng.optimize_batched(qn.Rosenbrock(), Th, max_iter=cap, ntrace=cap)
        ref = ob.optimize_batch(ob.rosenbrock(), Th, max_iter=cap, ntrace=cap)
        for k in ("x", "g", "trace", "fevals", "iters", "status"):
            assert np.array_equal(got[k], ref[k]), (cap, k)


def test_committed_oracle_vectors(eng, qn):
    """Against the committed fixtures (no oracle call): tests/golden/oracle_vectors.npz."""
    v = np.load(os.path.join(GOLDEN, "oracle_vectors.npz"))
    keys = sorted({k.rsplit("_", 1)[0] for k in v.files})
    for k in keys:
        order = int(k.split("_")[1][1:])
        got = eng.optimize_batched(qn.Rosenbrock(), v[k + "_theta0"], ls=qn.BackTracking(order=order), ntrace=5)
        ref = {f: v[k + "_" + f] for f in ("x", "g", "ell", "iters", "fevals", "status", "trace")}
        _check(got, ref)


def test_first_five_iterates_match_mpmath_fixture(eng, qn):
    """The GPU's first five iterates against the independent 60-digit restatement."""
    import json
    d = json.load(open(os.path.join(GOLDEN, "rosenbrock_mp_trace.json")))
    for case in d["cases"]:
        x0 = np.array([float.fromhex(h) for h in case["x0"]]).reshape(-1, 1)
        got = eng.optimize_batched(qn.Rosenbrock(), x0, ls=qn.BackTracking(order=case["order"]), ntrace=5)
        ref = np.array(case["trace"])
        scale = np.abs(ref).max(axis=1, keepdims=True)
        assert (np.abs(got["trace"][0] - ref) / scale).max() <= TOL_TRACE


def test_full_size_batch(eng, qn, ob):
    """BASELINE.json configs[1]: 2^16 starts, n = 60.  Every problem converges; a 2048-problem
    sample is compared with the oracle bit for bit; results do not depend on batch composition."""
    n, batch = 60, 1 << 16
    Theta = starts(n, batch)
    got = eng.optimize_batched(qn.Rosenbrock(), Theta)
    assert eng.last_launch()["block"] == 128  # at this size the tiles live in tensor memory (8 problems per SM)
    assert (got["status"] == qn.CONVERGED).all()
    assert np.abs(got["g"]).max() < TOL_G
    assert np.abs(got["x"] - 1).max() < 2e-7
    assert np.abs(got["ell"]).max() < 1e-14
    idx = np.arange(0, batch, 32)
    ref = ob.optimize_batch(ob.rosenbrock(), Theta[:, idx])
    sub = {k: (got[k][..., idx] if got[k] is not None else None) for k in got}
    _check(sub, ref)
    # size-independent property: a problem's result is the same alone, in a small batch, reversed
    alone = eng.optimize_batched(qn.Rosenbrock(), Theta[:, 4097:4098])
    assert np.array_equal(alone["x"][:, 0], got["x"][:, 4097])
    rev = eng.optimize_batched(qn.Rosenbrock(), Theta[:, 1023::-1])
    assert np.array_equal(rev["x"][:, ::-1], got["x"][:, :1024])
    assert np.array_equal(rev["iters"][::-1], got["iters"][:1024])
    print("2^16 batch: iterations mean %.1f median %d max %d; fevals/iter %.3f" % (
        got["iters"].mean(), np.median(got["iters"]), got["iters"].max(),
        got["fevals"].sum() / got["iters"].sum()))


def test_edge_cases(eng, qn, ob):
    R = qn.Rosenbrock()
    # empty batch
    r = eng.optimize_batched(R, np.empty((7, 0), order="F"))
    assert r["ell"].shape == (0,)
    # non-finite start -> NaN + status (QN.jl:255,291); finite neighbours unaffected
    Theta = starts(10, 6)
    Theta[3, 2] = np.inf
    Theta[0, 4] = np.nan
    got = eng.optimize_batched(R, Theta, ntrace=2)
    ref = ob.optimize_batch(ob.rosenbrock(), Theta, ntrace=2)
    _check(got, ref)
    assert got["status"][2] == qn.NONFINITE and np.isnan(got["ell"][2])
    assert got["status"][4] == qn.NONFINITE and got["status"][0] == qn.CONVERGED
    # iteration cap -> NaN (QN.jl:291); `gradient(state)` then shows the buffer the reference's
    # pointer swaps (QN.jl:286) leave in the canonical slot: odd cap -> latest ∇, even -> previous
    for n, cap in ((10, 3), (10, 4), (60, 5), (60, 6), (24, 1), (24, 2)):
        Th = starts(n, 4, key=cap)
        got = eng.optimize_batched(R, Th, max_iter=cap, ntrace=cap)
        ref = ob.optimize_batch(ob.rosenbrock(), Th, max_iter=cap, ntrace=cap)
        assert (got["status"] == qn.MAX_ITER).all() and np.isnan(got["ell"]).all() and (got["iters"] == cap).all()
        for k in ("x", "g", "trace", "fevals"):
            assert np.array_equal(got[k], ref[k]), (n, cap, k)
    # already at the optimum: no step taken
    got = eng.optimize_batched(R, np.ones((9, 3), order="F"))
    assert (got["iters"] == 0).all() and (got["ell"] == 0).all() and (got["status"] == 0).all()
    # huge starts: the non-finite halving loop (QN.jl:178-184) and resets are exercised
    big = starts(8, 64, key=5) * 1e80
    got = eng.optimize_batched(R, big)
    ref = ob.optimize_batch(ob.rosenbrock(), big)
    _check(got, ref)
    # tiny c1 / aggressive clamps, user BackTracking parameters pass through
    ls = qn.BackTracking(c1=0.3, rho_hi=0.9, rho_lo=0.5, iterations=3, order=3)
    lo = ob.backtracking(c1=0.3, rho_hi=0.9, rho_lo=0.5, iterations=3, order=3)
    got = eng.optimize_batched(R, starts(12, 64, key=9), ls=ls)
    ref = ob.optimize_batch(ob.rosenbrock(), starts(12, 64, key=9), ls=lo)
    _check(got, ref)
    assert (ref["status"] == ob.LINESEARCH_FAILED).any()  # the α = 0 exit (QN.jl:193,284)


def test_state_blocks_and_h_writeback(eng, qn, ob):
    """Results land in BFGSState layout (QN.jl:135,138): θ* at 0, ∇ at L; with QN_FLAG_WRITE_H the
    final B⁻¹ at 7L equals the oracle's bit for bit."""
    for n, flags in ((60, 0), (60, qn._lib.FLAG_FORCE_SMEM), (24, 0), (8, 0), (16, 0), (61, qn._lib.FLAG_FORCE_SMEM),
                     (128, 0), (128, qn._lib.FLAG_FORCE_SMEM), (159, 0)):
        batch = 16 if n < 100 else 3
        Theta = starts(n, batch, key=n + 1)
        sb = qn.BFGSStateBatch(n, batch)
        sb.data[:] = np.nan
        r = eng.optimize_batched(qn.Rosenbrock(), Theta, states=sb, flags=flags | qn._lib.FLAG_WRITE_H)
        assert np.array_equal(qn.optimum(sb).T, r["x"]) and np.array_equal(qn.gradient(sb).T, r["g"])
        for j in range(batch):
            o = ob.optimize(ob.rosenbrock(), Theta[:, j], L=sb.L)
            L = sb.L
            Ho = o["state"][7 * L:].reshape(n, L)[:, :n]
            Hg = sb.data[j, 7 * L:].reshape(n, L)[:, :n]
            assert np.array_equal(Hg, Ho)
            assert np.array_equal(sb.data[j, :n], o["x"]) and np.array_equal(sb.data[j, L:L + n], o["g"])


def test_reference_api_single_problem(qn, ob):
    """optimize!(state, obj, x) exactly as the reference's test uses it (runtests.jl:39-50)."""
    eps = np.finfo(float).eps
    for n in (2, 9, 24):
        state = qn.BFGSState(n)
        x = starts(n, 1, key=1000 + n)[:, 0]
        x_before = x.copy()
        ell = qn.optimize_(state, qn.Rosenbrock(), x)
        assert abs(ell) < 2 * eps
        assert np.all(np.isclose(qn.optimum(state), 1, rtol=np.sqrt(eps), atol=0))
        assert np.max(np.abs(qn.gradient(state))) < 1e-8
        assert np.array_equal(x, x_before)  # x is read-only (QN.jl:248)
        ell3 = qn.optimize_(state, qn.Rosenbrock(), x, qn.BackTracking[3]())
        assert abs(ell3) < eps
        assert np.max(np.abs(qn.gradient(state))) < 1e-8
        o = ob.optimize(ob.rosenbrock(), x, ls=ob.backtracking(order=3))
        assert ell3 == o["ell"] and np.array_equal(qn.optimum(state), o["x"]) and state.iters == o["iters"]
    # batched optimize!(states, obj, Θ)
    sb = qn.BFGSStateBatch(12, 32)
    ell = qn.optimize_(sb, qn.Rosenbrock(), starts(12, 32))
    assert ell.shape == (32,) and (np.abs(ell) < 2 * eps).all() and (sb.status == 0).all()


def test_unregistered_objective_is_an_error(eng, qn):
    class NotOnTheDevice:
        pass
    with pytest.raises(qn.UnregisteredObjectiveError):
        eng.optimize_batched(NotOnTheDevice(), starts(4, 2))
    import ctypes as C
    h = C.c_void_p()
    rc = eng._lib.qn_model_create(eng._ctx, 77, C.byref(h))
    assert rc == qn._lib.ERR_UNREGISTERED_OBJECTIVE and not h.value


def test_device_objectives_match_oracle(eng, qn, ob):
    """logdensity / ∂logdensity! (QN.jl:8-9) of the registered device objectives."""
    for n in (2, 7, 60, 61):
        Th = starts(n, 33, key=n) * 1.5
        ell, g = eng.logdensity_batched(qn.Rosenbrock(), Th, grad=True)
        for j in range(Th.shape[1]):
            v, gr = ob.dlogdensity(ob.rosenbrock(), Th[:, j])
            assert v == ell[j] and np.array_equal(gr, g[:, j])
            assert ob.logdensity(ob.rosenbrock(), Th[:, j]) == ell[j]
    x = starts(9, 1)[:, 0]
    gbuf = np.empty(9)
    assert qn.dlogdensity_(gbuf, qn.Rosenbrock(), x) == qn.logdensity(qn.Rosenbrock(), x)
    assert np.array_equal(gbuf, ob.dlogdensity(ob.rosenbrock(), x)[1])


@pytest.mark.parametrize("n", [36, 50, 60, 64])
def test_three_warp_speculative_kernel(eng, qn, ob, n, option):
    """Option tile_3w = 1 (qn_resident_tile3w.cuh): two helper warps own the tile, a third warp the
    vectors, and pass 1 of the next update is swept speculatively while the first trial is still being
    judged (discarded when another step is taken, B⁻¹ is reset or the solve ends).  Same bits as the
    oracle: traces, counts, both line-search orders, the final B⁻¹, iteration caps, ragged batches."""
    option("tile_tmem", 1)
    option("tile_3w", 1)
    R = qn.Rosenbrock()
    for batch in (1, 3, 23):
        Theta = starts(n, batch, key=8300 + n + batch)
        for order in (2, 3):
            got = eng.optimize_batched(R, Theta, ls=qn.BackTracking[order](), ntrace=5)
            assert eng.last_launch()["block"] == 384
            ref = ob.optimize_batch(ob.rosenbrock(), Theta, ls=ob.backtracking(order), ntrace=5)
            _check(got, ref)
    batch = 5
    Theta = starts(n, batch, key=8400 + n)
    sb = qn.BFGSStateBatch(n, batch)
    eng.optimize_batched(R, Theta, states=sb, flags=qn._lib.FLAG_WRITE_H)
    L = sb.L
    for j in range(batch):
        o = ob.optimize(ob.rosenbrock(), Theta[:, j], L=L)
        assert np.array_equal(sb.data[j, 7 * L:].reshape(n, L)[:, :n], o["state"][7 * L:].reshape(n, L)[:, :n])
        assert np.array_equal(sb.data[j, :n], o["x"]) and np.array_equal(sb.data[j, L:L + n], o["g"])
    for cap in (1, 2, 5, 6):
        Th = starts(n, 4, key=cap)
        if cap == 5:
            Th[3, 2] = np.nan
        got = eng.optimize_batched(R, Th, max_iter=cap, ntrace=cap)
        ref = ob.optimize_batch(ob.rosenbrock(), Th, max_iter=cap, ntrace=cap)
        for k in ("x", "g", "trace", "ell"):
            assert np.array_equal(got[k], ref[k], equal_nan=True), (n, cap, k)
        for k in ("iters", "fevals", "status"):
            assert np.array_equal(got[k], ref[k]), (n, cap, k)


@pytest.mark.parametrize("n", [36, 52, 60])
def test_leader_placement_is_not_numerics(eng, qn, ob, n, option):
    """Option lead_parity picks WHICH warp of a leader/helper pair leads (0: the same in every CTA,
    the default; -1 / d: alternating with blockIdx) — scheduling only: the bits equal the oracle's for
    every placement, in tensor memory and in the all-register kernel (which runs without the split)."""
    Theta = starts(n, 1500, key=8600 + n)
    idx = np.arange(0, 1500, 47)
    ref = ob.optimize_batch(ob.rosenbrock(), Theta[:, idx])
    for tm in (1, 0):
        option("tile_tmem", tm)
        for parity in (0, -1, 1, 3):
            option("lead_parity", parity)
            got = eng.optimize_batched(qn.Rosenbrock(), Theta)
            assert eng.last_launch()["block"] == (128 if tm else 64)
            for k in ("x", "g", "ell"):
                assert np.array_equal(got[k][..., idx], ref[k]), (tm, parity, k)
            for k in ("iters", "fevals", "status"):
                assert np.array_equal(got[k][idx], ref[k]), (tm, parity, k)

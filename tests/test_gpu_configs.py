"""The BASELINE.json configurations at their FULL sizes and the
device-pointer entry — the cases the round-1 review found untested:

  * config 5: two-pass (the reference's order) at n = 8192 — first iterations bit for bit against the
    oracle, and ONE full solve to convergence with its iteration count and time per solve reported;
  * `qn_bfgs_optimize_batched_device_f64` with device-resident BFGSState blocks + QN_FLAG_WRITE_H
    (the PtrBFGSState analogue, QN.jl:117-121) against the oracle;
  * config 3: the full 2^15 dense-Gaussian and logistic batches, a seeded sample of them against
    the oracle plus size-independent properties of every problem;
  * config 2: the full 2^16 Rosenbrock batch through the host API and through the multi-GPU entry
    (all visible GPUs), a sample against the oracle plus the known answer for every problem.
"""
import json
import os
import time

import numpy as np
import pytest

from conftest import ROOT, starts
from problems import gaussian_problem, logistic_problem

pytestmark = pytest.mark.gpu


def _report(name, obj):
    """Measured facts of a full-size case, kept beside the test run (copied to profiles/ by hand)."""
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, name), "w") as f:
        json.dump(obj, f, indent=1)
    print(name, json.dumps(obj))


# ------------------------------------------------------------------ config 5: n = 8192, two-pass
def test_n8192_two_pass_first_iterations_bit_exact(eng, qn, ob):
    n, iters = 8192, 6
    Theta = starts(n, 1, key=8192)
    got = eng.optimize_batched(qn.Rosenbrock(), Theta, max_iter=iters, ntrace=iters)
    assert eng.last_launch()["path"] == qn._lib.PATH_STREAM
    ref = ob.optimize_batch(ob.rosenbrock(), Theta, max_iter=iters, ntrace=iters)
    for k in ("x", "g", "ell", "trace"):
        assert np.array_equal(got[k], ref[k], equal_nan=True), k
    for k in ("iters", "fevals", "status"):
        assert np.array_equal(got[k], ref[k]), k


def test_n8192_full_solve_to_convergence(eng, qn):
    """SURVEY §8(d) config 5 asks for the time per SOLVE as well as per iteration.  The known answer
    (runtests.jl:42-45: |ℓ*| tiny, θ* = 1, max|∇| < 1e-8) is path-independent, so it is the gate."""
    n = 8192
    Theta = starts(n, 1, key=8193)
    eng.optimize_batched(qn.Rosenbrock(), Theta, max_iter=8)  # warm-up: allocation, graph capture
    t = time.perf_counter()
    got = eng.optimize_batched(qn.Rosenbrock(), Theta)
    dt = time.perf_counter() - t
    assert got["status"][0] == qn.CONVERGED, got["status"]
    assert np.abs(got["g"]).max() < 1e-8
    err = float(np.abs(got["x"] - 1.0).max())
    assert err < 1e-7
    assert abs(got["ell"][0]) < 1e-10
    _report("r2_n8192_full_solve.json", dict(
        n=n, mode="two-pass (reference order)", iterations=int(got["iters"][0]), fevals=int(got["fevals"][0]),
        wall_s_per_solve=dt, us_per_iteration=dt / max(int(got["iters"][0]), 1) * 1e6,
        max_abs_grad=float(np.abs(got["g"]).max()), max_abs_theta_minus_1=err, ell=float(got["ell"][0])))


# ------------------------------------------------------------------ device-pointer entry (PtrBFGSState analogue)
@pytest.mark.parametrize("n", [60, 128, 24])
def test_device_entry_with_device_states_and_h(eng, qn, ob, n):
    """Every pointer a device pointer, results straight into device-resident BFGSState blocks with
    the final B⁻¹ (QN_FLAG_WRITE_H), on a caller-owned stream — against the oracle's state block."""
    import torch
    dev = torch.device("cuda", 0)
    batch = 37
    L = qn._lib.load().qn_state_ld(n, 8)
    stride = qn._lib.load().qn_state_len(n, L) + 8  # a caller's own, larger stride
    Theta = starts(n, batch, key=4242 + n)
    th_d = torch.from_numpy(np.ascontiguousarray(Theta.T)).to(dev)  # (batch, n) C-order = n×batch column-major
    st_d = torch.full((batch, stride), float("nan"), dtype=torch.float64, device=dev)
    ell_d = torch.empty(batch, dtype=torch.float64, device=dev)
    it_d = torch.empty(batch, dtype=torch.int32, device=dev)
    fe_d = torch.empty(batch, dtype=torch.int32, device=dev)
    s_d = torch.empty(batch, dtype=torch.int32, device=dev)
    stream = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(stream):
        eng.optimize_batched_device(qn.Rosenbrock(), n, batch, th_d.data_ptr(), states_ptr=st_d.data_ptr(),
                                    state_stride=stride, ld=L, ell_ptr=ell_d.data_ptr(), iters_ptr=it_d.data_ptr(),
                                    fevals_ptr=fe_d.data_ptr(), status_ptr=s_d.data_ptr(),
                                    flags=qn._lib.FLAG_WRITE_H, stream=stream.cuda_stream)
    stream.synchronize()
    st = st_d.cpu().numpy()
    for j in range(batch):
        o = ob.optimize(ob.rosenbrock(), Theta[:, j], L=L)
        assert np.array_equal(st[j, :n], o["x"]) and np.array_equal(st[j, L:L + n], o["g"]), j
        assert np.array_equal(st[j, 7 * L:7 * L + n * L].reshape(n, L)[:, :n],
                              o["state"][7 * L:].reshape(n, L)[:, :n]), j
        assert ell_d[j].item() == o["ell"] and it_d[j].item() == o["iters"] and fe_d[j].item() == o["fevals"]
        assert s_d[j].item() == o["status"]
    # what the entry must leave alone: the scratch fields 3..7 and the slack behind each block
    assert np.isnan(st[:, 2 * L:7 * L]).all() and np.isnan(st[:, 7 * L + n * L:]).all()


def test_device_entry_rejects_bad_strides(eng, qn):
    import torch
    t = torch.zeros(60 * 4, dtype=torch.float64, device="cuda:0")
    with pytest.raises(qn.QNError):
        eng.optimize_batched_device(qn.Rosenbrock(), 60, 4, t.data_ptr(), theta_stride=59, x_ptr=t.data_ptr())
    with pytest.raises(qn.QNError):
        eng.optimize_batched_device(qn.Rosenbrock(), 60, 4, t.data_ptr(), states_ptr=t.data_ptr(), state_stride=100, ld=64)


# ------------------------------------------------------------------ config 3 at its full batch
def _sample(batch, k, key):
    return np.sort(np.random.Generator(np.random.Philox(key=key)).choice(batch, size=k, replace=False))


def test_config3_gaussian_full_batch(eng, qn, ob):
    n, batch = 128, 1 << 15
    mu, lam = gaussian_problem(n)
    Theta = starts(n, batch, key=333)
    got = eng.optimize_batched(qn.DenseGaussian(mu, lam), Theta)
    # every problem: the maximiser of a Gaussian is μ, ℓ* = 0
    assert (got["status"] == qn.CONVERGED).all()
    assert np.abs(got["g"]).max() < 1e-8
    assert np.abs(got["x"] - mu[:, None]).max() < 1e-7
    assert np.abs(got["ell"]).max() < 1e-14
    # a seeded sample: bit for bit against the oracle
    idx = _sample(batch, 48, 334)
    ref = ob.optimize_batch(ob.OracleModel(ob.DENSE_GAUSSIAN, mu=mu, lam=lam), Theta[:, idx])
    for k in ("x", "g", "ell"):
        assert np.array_equal(got[k][..., idx], ref[k]), k
    for k in ("iters", "fevals", "status"):
        assert np.array_equal(got[k][idx], ref[k]), k


def test_config3_logistic_full_batch(eng, qn, ob):
    n, m, batch = 128, 1024, 1 << 15
    X, y, lam = logistic_problem(m, n)
    Theta = starts(n, batch, key=335)
    got = eng.optimize_batched(qn.LogisticRegression(X, y, lam), Theta)
    assert set(np.unique(got["status"])) <= {qn.CONVERGED, qn.LINESEARCH_FAILED}
    conv = got["status"] == qn.CONVERGED
    assert conv.mean() >= 0.85
    assert np.abs(got["g"][:, conv]).max() < 1e-8
    assert np.abs(got["g"]).max() < 1e-6  # line-search failures stop at the rounding floor of ℓ
    # strictly concave: ONE maximiser for every start
    xs = got["x"][:, conv]
    assert np.abs(xs - xs[:, :1]).max() <= 2e-8
    idx = _sample(batch, 24, 336)
    ref = ob.optimize_batch(ob.OracleModel(ob.LOGISTIC, X=X, y=y, lam=lam), Theta[:, idx])
    both = conv[idx] & (ref["status"] == ob.CONVERGED)
    assert both.mean() >= 0.7
    assert np.abs(got["x"][:, idx][:, both] - ref["x"][:, both]).max() <= 1e-8  # tolerance: CUDA vs glibc exp/log1p
    assert np.abs(got["ell"][idx][both] - ref["ell"][both]).max() <= 1e-12 * np.abs(ref["ell"][both]).max()


# ------------------------------------------------------------------ config 2 at its full batch
def test_config2_full_batch_host_and_multi_entry(eng, qn, ob):
    n, batch = 60, 1 << 16
    Theta = starts(n, batch)
    got = eng.optimize_batched(qn.Rosenbrock(), Theta)
    assert eng.last_launch()["block"] in (128, 384)  # B⁻¹ tiles in tensor memory
    # the reference's own assertions (runtests.jl:42-45, README.md:59-75) for every problem
    assert (got["status"] == qn.CONVERGED).all()
    assert np.abs(got["g"]).max() < 1e-8
    assert np.abs(got["x"] - 1.0).max() < 1e-7
    assert np.abs(got["ell"]).max() < 1e-15
    idx = _sample(batch, 96, 61)
    ref = ob.optimize_batch(ob.rosenbrock(), Theta[:, idx])
    for k in ("x", "g", "ell"):
        assert np.array_equal(got[k][..., idx], ref[k]), k
    for k in ("iters", "fevals", "status"):
        assert np.array_equal(got[k][idx], ref[k]), k
    # the same batch sharded over every visible GPU from one process: identical bits, problem by problem
    G = qn._lib.load().qn_device_count()
    engines = [eng] + [qn.Engine(d) for d in range(1, G)]
    multi = qn.optimize_batched_multi(engines, qn.Rosenbrock(), Theta)
    for k in ("x", "g", "ell", "iters", "fevals", "status"):
        assert np.array_equal(multi[k], got[k], equal_nan=True), k
    for e in engines[1:]:
        e.close()


# ------------------------------------------------------------------ host entries: pipelined column ranges
@pytest.mark.parametrize("n,batch", [(8, 5000), (60, 4099), (24, 16411)])
def test_host_entry_ranges_same_bits(eng, qn, option, n, batch):
    """Option "host_chunks": the host-pointer entry cuts a batch into column ranges pipelined over
    the context's streams (copies under kernels, the next kernel filling the tail of the previous).
    One range, the automatic choice and odd range counts return the same bytes, BFGSState blocks and
    B⁻¹ (QN.jl:135-142) included."""
    Theta = starts(n, batch)
    ref_states = qn.BFGSStateBatch(n, batch)
    ref_states.data[...] = 0.0  # the fields the solve does not return keep the caller's bytes
    option("host_chunks", 1)
    ref = eng.optimize_batched(qn.Rosenbrock(), Theta, states=ref_states, flags=qn._lib.FLAG_WRITE_H)
    assert eng.last_launch()["kernel_launches"] == 1
    for chunks in (-1, 3, 8):
        option("host_chunks", chunks)
        st = qn.BFGSStateBatch(n, batch)
        st.data[...] = 0.0
        got = eng.optimize_batched(qn.Rosenbrock(), Theta, states=st, flags=qn._lib.FLAG_WRITE_H)
        want = {(-1): min(8, batch // 2048)}.get(chunks, chunks)
        assert eng.last_launch()["kernel_launches"] == want
        for k in ("x", "g", "ell", "iters", "fevals", "status"):
            assert np.array_equal(got[k], ref[k], equal_nan=True), (chunks, k)
        assert np.array_equal(st.data, ref_states.data, equal_nan=True), chunks
    # a recorded trace keeps the batch in one launch
    option("host_chunks", 8)
    tr = eng.optimize_batched(qn.Rosenbrock(), Theta[:, :4200], ntrace=3)
    assert eng.last_launch()["kernel_launches"] == 1
    one = eng.optimize_batched(qn.Rosenbrock(), Theta[:, :4200])
    assert np.array_equal(tr["x"], one["x"])


def test_multi_entry_ranges_same_bits(eng, qn, option):
    """The multi-GPU entry pipelines every device's share the same way; ragged shares and range
    counts that do not divide them give the bytes of the one-range call."""
    n, batch = 60, 9001
    Theta = starts(n, batch)
    G = qn._lib.load().qn_device_count()
    engines = [eng] + [qn.Engine(d) for d in range(1, G)]
    try:
        for e in engines:
            e.set_option("host_chunks", 1)
        ref = qn.optimize_batched_multi(engines, qn.Rosenbrock(), Theta)
        for chunks in (-1, 3, 8):
            for e in engines:
                e.set_option("host_chunks", chunks)
            got = qn.optimize_batched_multi(engines, qn.Rosenbrock(), Theta)
            for k in ("x", "g", "ell", "iters", "fevals", "status"):
                assert np.array_equal(got[k], ref[k], equal_nan=True), (chunks, k)
        assert (ref["status"] == qn.CONVERGED).all()
    finally:
        eng.set_option("host_chunks", -1)
        for e in engines[1:]:
            e.close()

"""Run-to-run determinism under full load.  compute-sanitizer is not available on the GPU pool, so
a shared-memory race or a missing barrier would have to show up as a result that changes between
identical launches with every SM busy (and different problems co-resident each time, because the
persistent grids pull problems from an atomic work queue).  Parity with the oracle on samples of the
same batches is asserted in test_gpu_parity.py; here the whole batch must repeat bit for bit."""
import numpy as np
import pytest

from conftest import starts

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,batch", [(8, 1 << 17), (16, 1 << 16), (32, 1 << 15), (60, 1 << 15), (64, 1 << 13),
                                     (100, 1 << 10), (128, 1 << 11), (160, 1 << 9)])
def test_repeated_launches_agree_bitwise(eng, qn, ob, n, batch):
    Theta = starts(n, batch, key=4242 + n)
    first = eng.optimize_batched(qn.Rosenbrock(), Theta)
    for _ in range(2):
        again = eng.optimize_batched(qn.Rosenbrock(), Theta)
        for k in ("x", "g", "ell", "iters", "fevals", "status"):
            assert np.array_equal(first[k], again[k], equal_nan=True), k
    # and a sample of the batch against the oracle
    sel = np.linspace(0, batch - 1, 24).astype(int)
    ref = ob.optimize_batch(ob.rosenbrock(), np.asfortranarray(Theta[:, sel]))
    for k in ("x", "g", "ell", "iters", "fevals", "status"):
        assert np.array_equal(first[k][..., sel], ref[k], equal_nan=True), k


def test_objectives_repeat_bitwise(eng, qn):
    from problems import gaussian_problem, logistic_problem
    mu, lam = gaussian_problem(128)
    X, y, l2 = logistic_problem(1024, 128)
    for obj, batch in ((qn.DenseGaussian(mu, lam), 2048), (qn.LogisticRegression(X, y, l2), 300)):
        Theta = starts(128, batch, key=99)
        a = eng.optimize_batched(obj, Theta)
        b = eng.optimize_batched(obj, Theta)
        for k in ("x", "g", "ell", "iters", "fevals", "status"):
            assert np.array_equal(a[k], b[k], equal_nan=True), k


def test_distinct_contexts_are_independent_across_host_threads(qn, ob):
    """calls on distinct contexts are independent — the analogue of calling the
    reference from Threads.@threads with distinct states.  Four host threads, four contexts on the
    same GPU, different problem shapes at once; every result equals the oracle's."""
    import threading
    shapes = [(60, 400), (8, 3000), (128, 40), (100, 60)]
    engines = [qn.Engine(0) for _ in shapes]
    out = [None] * len(shapes)

    def work(k):
        n, batch = shapes[k]
        Theta = starts(n, batch, key=600 + k)
        for _ in range(3):
            out[k] = (Theta, engines[k].optimize_batched(qn.Rosenbrock(), Theta))

    threads = [threading.Thread(target=work, args=(k,)) for k in range(len(shapes))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for k, (n, batch) in enumerate(shapes):
        Theta, got = out[k]
        sel = np.linspace(0, batch - 1, 16).astype(int)
        ref = ob.optimize_batch(ob.rosenbrock(), np.asfortranarray(Theta[:, sel]))
        for key in ("x", "g", "ell", "iters", "fevals", "status"):
            assert np.array_equal(got[key][..., sel], ref[key], equal_nan=True), (n, key)
    for e in engines:
        e.close()


def test_two_threads_on_one_context(eng, qn, ob):
    """Two host threads calling the host entry of ONE context at the same time (ctypes releases the
    GIL): the context's mutex is held for the whole call, so neither can see — or free — the other's
    staging (round-1 advisor finding).  Different sizes per thread so that a mixed-up staging buffer
    could not go unnoticed; each result must be the oracle's."""
    import threading
    jobs = [(60, 700, 1), (24, 1500, 2), (60, 64, 3), (100, 90, 4)]
    out = {}

    def work(tid):
        res = []
        for rep in range(3):
            for n, batch, key in jobs[tid::2]:
                Theta = starts(n, batch, key=key)
                res.append((n, key, Theta, eng.optimize_batched(qn.Rosenbrock(), Theta)))
        out[tid] = res

    ts = [threading.Thread(target=work, args=(t,)) for t in range(2)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    for tid in range(2):
        assert len(out[tid]) == 6
        for n, key, Theta, got in out[tid]:
            sel = np.linspace(0, Theta.shape[1] - 1, 16).astype(int)
            ref = ob.optimize_batch(ob.rosenbrock(), np.asfortranarray(Theta[:, sel]))
            for k in ("x", "g", "ell", "iters", "fevals", "status"):
                assert np.array_equal(got[k][..., sel], ref[k], equal_nan=True), (n, key, k)

"""Host-side mirror of the reference interface (no GPU needed)."""
import numpy as np
import pytest


def test_backtracking_defaults(qn):
    # QN.jl:79-80
    ls = qn.BackTracking()
    assert (ls.c1, ls.rho_hi, ls.rho_lo, ls.iterations, ls.order) == (1e-4, 0.5, 0.1, 1000, 2)
    ls3 = qn.BackTracking[3]()
    assert ls3.order == 3 and ls3.iterations == 1000 and isinstance(ls3, qn.BackTracking)
    assert qn.BackTracking(order=3).order == 3
    with pytest.raises(ValueError):
        qn.BackTracking(order=4)
    c = ls3._c()
    assert c.order == 3 and c.c1 == 1e-4


def test_bfgsstate_layout_and_views(qn):
    # QN.jl:84-104, 135-142, 149-150
    st = qn.BFGSState(60)
    assert st.L == 64 and st.data.nbytes == 34304
    st.data[:] = np.arange(st.data.size)
    assert np.array_equal(qn.optimum(st), np.arange(60))
    assert np.array_equal(qn.gradient(st), 64 + np.arange(60))
    B = st.B_inv
    assert B.shape == (60, 60)
    assert B[3, 5] == 7 * 64 + 3 + 64 * 5  # H[c, r] at 7L + c + L r
    assert qn.BFGSState(60, native_w=4).L == 60
    sb = qn.BFGSStateBatch(24, 10)
    assert sb.data.shape == (10, 7 * 24 + 24 * 24) and len(sb) == 10
    assert qn.optimum(sb).shape == (10, 24) and qn.gradient(sb).shape == (10, 24)
    assert sb[3].data.base is not None  # a view, not a copy


def test_unregistered_objective_raises_without_touching_the_gpu(qn):
    class MyJuliaStyleModel:  # has no device implementation
        pass

    class FakeEngine(qn.Engine):
        def __init__(self):
            self._models, self._keep, self._ctx = {}, {}, None

    with pytest.raises(qn.UnregisteredObjectiveError) as e:
        FakeEngine().model_handle(MyJuliaStyleModel())
    assert "no CPU fallback" in str(e.value)


def test_sharding_roundtrip(qn):
    from qn_b200 import sharding
    batch, n = 37, 5
    Theta = np.arange(n * batch, dtype=np.float64).reshape(n, batch, order="F")
    for world in (1, 2, 3, 8):
        parts = [sharding.scatter(Theta, r, world) for r in range(world)]
        assert sum(p.shape[1] for p in parts) == batch
        assert np.array_equal(sharding.gather(parts, batch), Theta)
        idx = np.concatenate([sharding.shard_columns(batch, r, world) for r in range(world)])
        assert sorted(idx.tolist()) == list(range(batch))

"""Host-side mirror of the reference's interface for the BFGS hot path.

Reference (QN.jl = /root/reference/src/QuasiNewtonMethods.jl)      here
------------------------------------------------------------      -------------------------
BFGSState{P}(undef)                       QN.jl:84-105            BFGSState(P)
optimize!(state, obj, x, ls, tol)         QN.jl:237               optimize_(state, obj, x, ls, tol)
optimum(state) / gradient(state)          QN.jl:149-150           optimum(state) / gradient(state)
BackTracking() / BackTracking{3}()        QN.jl:72-80             BackTracking() / BackTracking[3]()
logdensity(obj, θ) / ∂logdensity!(∇,obj,θ) QN.jl:8-9              logdensity(obj, θ) / dlogdensity_(∇, obj, θ)
(new) optimize!(states, obj, Θ)           BASELINE.json           optimize_(BFGSStateBatch, obj, Θ)

Same names, argument meaning and error behaviour: failures of a problem are
in-band (NaN return, QN.jl:291); an objective type without a registered device
implementation raises UnregisteredObjectiveError, the MethodError analogue
(QN.jl:254) — there is no CPU fallback.  Everything numeric happens in
libqn_cuda.so; this file only marshals pointers.
"""
import ctypes as C
import threading

import numpy as np

from . import _lib
from ._lib import (CONVERGED, LINESEARCH_FAILED, MAX_ITER, NONFINITE, QNError,
                   UnregisteredObjectiveError)

__all__ = [
    "BFGSState", "BFGSStateBatch", "BackTracking", "Rosenbrock", "DenseGaussian",
    "LogisticRegression", "UserObjective", "optimize_", "optimize", "optimum", "gradient", "logdensity",
    "dlogdensity_", "Engine", "default_engine", "optimize_rowsharded", "optimize_batched_multi", "QNError", "UnregisteredObjectiveError",
    "CONVERGED", "NONFINITE", "LINESEARCH_FAILED", "MAX_ITER",
]

MAX_ITER_DEFAULT = 10_000  # N = 10_000, QN.jl:250


def _ptr(a):
    return None if a is None else a.ctypes.data


# ------------------------------------------------------------------ BackTracking
class BackTracking:
    """BackTracking{O}(c₁=1e-4, ρₕ=0.5, ρₗ=0.1, iterations=1000), QN.jl:72-80.

    `BackTracking()` is order 2 (QN.jl:80); `BackTracking[3]()` is the
    reference's `BackTracking{3}()` (QN.jl:79).
    """
    order = 2

    def __init__(self, c1=1e-4, rho_hi=0.5, rho_lo=0.1, iterations=1000, order=None):
        self.c1, self.rho_hi, self.rho_lo, self.iterations = float(c1), float(rho_hi), float(rho_lo), int(iterations)
        if order is not None:
            self.order = int(order)
        if self.order not in (2, 3):
            raise ValueError("BackTracking order must be 2 or 3")

    def __class_getitem__(cls, order):
        return type("BackTracking{%d}" % order, (cls,), {"order": int(order)})

    def _c(self):
        return _lib.Backtracking(self.c1, self.rho_hi, self.rho_lo, self.iterations, self.order, 0)

    def __repr__(self):
        return "BackTracking{%d}(%g, %g, %g, %d)" % (self.order, self.c1, self.rho_hi, self.rho_lo, self.iterations)


# ------------------------------------------------------------------ objectives
class Rosenbrock:
    """The reference's example objective, test/runtests.jl:4-33 (kind QN_OBJ_ROSENBROCK)."""
    kind = _lib.OBJ_ROSENBROCK

    def _key(self):
        return ("rosenbrock",)

    def _register(self, engine):
        h = C.c_void_p()
        _lib.check(_lib.load().qn_model_create(engine._ctx, self.kind, C.byref(h)))
        return h


class DenseGaussian:
    """ℓ(θ) = -½ (θ-μ)ᵀ Λ (θ-μ).  Build-defined."""
    kind = _lib.OBJ_DENSE_GAUSSIAN

    def __init__(self, mu, lam):
        self.mu = np.ascontiguousarray(mu, dtype=np.float64)
        self.lam = np.asfortranarray(lam, dtype=np.float64)
        n = self.mu.size
        if self.lam.shape != (n, n):
            raise ValueError("Lambda must be n x n")

    def _key(self):
        return ("gaussian", id(self))

    def _register(self, engine):
        h = C.c_void_p()
        _lib.check(_lib.load().qn_model_dense_gaussian(engine._ctx, self.mu.size, _ptr(self.mu),
                                                       _ptr(self.lam), C.byref(h)))
        return h


class LogisticRegression:
    """ℓ(θ) = Σᵢ [yᵢ zᵢ - log(1+e^{zᵢ})] - ½ λ ‖θ‖², z = Xθ.  Build-defined."""
    kind = _lib.OBJ_LOGISTIC

    def __init__(self, X, y, lam=1.0):
        self.X = np.asfortranarray(X, dtype=np.float64)
        self.y = np.ascontiguousarray(y, dtype=np.float64)
        self.lam = float(lam)
        if self.X.shape[0] != self.y.size:
            raise ValueError("X must be m x n with m = len(y)")

    def _key(self):
        return ("logistic", id(self))

    def _register(self, engine):
        h = C.c_void_p()
        m, n = self.X.shape
        _lib.check(_lib.load().qn_model_logistic(engine._ctx, m, n, _ptr(self.X), _ptr(self.y),
                                                 self.lam, C.byref(h)))
        return h


class UserObjective:
    """The build's plug-in device objective (include/qn_objective.cuh, QN_OBJ_USER): `params` is the
    opaque parameter blob (a numpy array) the objective's `make` interprets; n = 0 accepts any n.
    Needs a library built with QN_USER_OBJECTIVE, selected with QN_LIBRARY_PATH; the stock library
    answers with UnregisteredObjectiveError (no CPU fallback)."""
    kind = _lib.OBJ_USER

    def __init__(self, params, n=0):
        self.params = np.ascontiguousarray(params)
        self.n = int(n)

    def _key(self):
        return ("user", id(self))

    def _register(self, engine):
        h = C.c_void_p()
        _lib.check(_lib.load().qn_model_user(engine._ctx, self.n, _ptr(self.params), self.params.nbytes,
                                             C.byref(h)))
        return h


# ------------------------------------------------------------------ states
class BFGSState:
    """BFGSState{P}(undef): one flat block of fp64 (QN.jl:84-92).

    Field order x_old, ∇_new, x_new, ∇_old, y, s, B⁻¹y (L each), then B⁻¹ (P×P,
    column stride L); L = align(P, pick_vector_width(P, Float64)) (QN.jl:96-99)
    for a host with `native_w` fp64 SIMD lanes.  Contents start uninitialised.
    """

    def __init__(self, P, native_w=8, L=None):
        self.P = int(P)
        lib = _lib.load()
        self.L = int(L) if L is not None else lib.qn_state_ld(self.P, native_w)
        self.data = np.empty(lib.qn_state_len(self.P, self.L), dtype=np.float64)
        self.iters = self.fevals = self.status = None

    @property
    def B_inv(self):
        """ref_B⁻¹(state): P×P view with column stride L (QN.jl:142)."""
        return self.data[7 * self.L:].reshape(self.P, self.L).T[:self.P, :]


class BFGSStateBatch:
    """`batch` BFGSState blocks in one array (block j = data[j]); the batched states argument."""

    def __init__(self, P, batch, native_w=8, L=None):
        self.P, self.batch = int(P), int(batch)
        lib = _lib.load()
        self.L = int(L) if L is not None else lib.qn_state_ld(self.P, native_w)
        self.stride = int(lib.qn_state_len(self.P, self.L))
        self.data = np.empty((self.batch, self.stride), dtype=np.float64)
        self.ell = self.iters = self.fevals = self.status = None

    def __len__(self):
        return self.batch

    def __getitem__(self, j):
        s = BFGSState.__new__(BFGSState)
        s.P, s.L, s.data = self.P, self.L, self.data[j]
        s.iters = None if self.iters is None else int(self.iters[j])
        s.fevals = None if self.fevals is None else int(self.fevals[j])
        s.status = None if self.status is None else int(self.status[j])
        return s


def optimum(state):
    """optimum(state) = ref_x_old(state): first P elements (QN.jl:135,149)."""
    if isinstance(state, BFGSStateBatch):
        return state.data[:, :state.P]
    return state.data[:state.P]


def gradient(state):
    """gradient(state) = ref_∇_new(state): elements L..L+P-1 (QN.jl:138,150)."""
    if isinstance(state, BFGSStateBatch):
        return state.data[:, state.L:state.L + state.P]
    return state.data[state.L:state.L + state.P]


# ------------------------------------------------------------------ engine
class Engine:
    """One qn_context (one GPU).  Owns the device-side registrations of objectives."""

    def __init__(self, device=0):
        self._lib = _lib.load()
        self._ctx = C.c_void_p()
        _lib.check(self._lib.qn_create(C.byref(self._ctx), device))
        self.device = device
        self._models = {}
        self._keep = {}

    def close(self):
        if getattr(self, "_ctx", None):
            for h in self._models.values():
                self._lib.qn_model_free(h)
            self._models.clear()
            self._lib.qn_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- objective registry: replaces dispatch on typeof(obj) (QN.jl:8-9)
    def model_handle(self, obj):
        if not hasattr(obj, "_register") or not hasattr(obj, "kind"):
            raise UnregisteredObjectiveError(
                _lib.ERR_UNREGISTERED_OBJECTIVE,
                "no device implementation registered for objective type %s "
                "(registered: Rosenbrock, DenseGaussian, LogisticRegression); "
                "there is no CPU fallback" % type(obj).__name__)
        key = obj._key()
        if key not in self._models:
            self._models[key] = obj._register(self)
            self._keep[key] = obj
        return self._models[key]

    def device_info(self):
        name = C.create_string_buffer(256)
        sm, smem, khz = C.c_int32(), C.c_int32(), C.c_int32()
        _lib.check(self._lib.qn_device_info(self._ctx, name, 256, C.byref(sm), C.byref(smem), C.byref(khz)))
        return dict(name=name.value.decode(), sm_count=sm.value, smem_per_block=smem.value, clock_khz=khz.value)

    def last_launch(self):
        v = [C.c_int32() for _ in range(5)]
        _lib.check(self._lib.qn_last_launch_info(self._ctx, *[C.byref(x) for x in v]))
        return dict(zip(("path", "grid", "block", "smem_bytes", "kernel_launches"), [x.value for x in v]))

    def set_option(self, key, value):
        """qn_set_option: tuning / ablation switch of this context (include/qn_cuda.h lists the keys).
        A one-character string stands for its character code ("pair_rows": 'h', 't', ...)."""
        if isinstance(value, str):
            value = ord(value[0]) if value else 0
        _lib.check(self._lib.qn_set_option(self._ctx, key.encode(), int(value)))

    def get_option(self, key):
        v = C.c_int64()
        _lib.check(self._lib.qn_get_option(self._ctx, key.encode(), C.byref(v)))
        return v.value

    def tail_profile(self):
        """Tail of the last tile-kernel launch made with option record_exit = 1."""
        span, idle, slots = C.c_double(), C.c_double(), C.c_int32()
        _lib.check(self._lib.qn_last_tail_profile(self._ctx, C.byref(span), C.byref(idle), C.byref(slots)))
        return dict(span_ms=span.value, idle_fraction=idle.value, slots=slots.value)

    def microbench(self):
        out = {}
        for name in ("dfma", "lds", "hbm"):
            v = C.c_double()
            _lib.check(getattr(self._lib, "qn_microbench_" + name)(self._ctx, C.byref(v)))
            out[name] = v.value
        return out

    # -- batched solve, host arrays
    def optimize_batched(self, obj, Theta, ls=None, tol=1e-8, max_iter=MAX_ITER_DEFAULT,
                         states=None, ntrace=0, flags=0, want_xg=True):
        """Solve one problem per COLUMN of Theta (n×batch).  Returns a dict of arrays."""
        Theta = np.asfortranarray(Theta, dtype=np.float64)
        if Theta.ndim != 2:
            raise ValueError("Theta must be n x batch")
        n, batch = Theta.shape
        ls = ls or BackTracking()
        lsc = ls._c()
        h = self.model_handle(obj)
        x = np.empty((n, batch), order="F") if want_xg else None
        g = np.empty((n, batch), order="F") if want_xg else None
        ell = np.empty(batch)
        iters = np.empty(batch, dtype=np.int32)
        fevals = np.empty(batch, dtype=np.int32)
        status = np.empty(batch, dtype=np.int32)
        trace = np.full((batch, ntrace, n), np.nan) if ntrace else None
        sp, sstride, sld = None, 0, 0
        if states is not None:
            if states.P != n or len(states) != batch:
                raise ValueError("states do not match Theta")
            sp, sstride, sld = _ptr(states.data), states.stride, states.L
        _lib.check(self._lib.qn_bfgs_optimize_batched_f64(
            self._ctx, h, n, batch, _ptr(Theta), C.byref(lsc), tol, max_iter, sp, sstride, sld,
            _ptr(x), _ptr(g), _ptr(ell), _ptr(iters), _ptr(fevals), _ptr(status), ntrace,
            _ptr(trace), flags))
        if states is not None:
            states.ell, states.iters, states.fevals, states.status = ell, iters, fevals, status
        return dict(x=x, g=g, ell=ell, iters=iters, fevals=fevals, status=status, trace=trace)

    # -- batched solve, device pointers (torch / CUDA.jl-style zero-copy callers)
    def optimize_batched_device(self, obj, n, batch, theta_ptr, ls=None, tol=1e-8,
                                max_iter=MAX_ITER_DEFAULT, theta_stride=None, states_ptr=None,
                                state_stride=0, ld=0, x_ptr=None, g_ptr=None, ell_ptr=None,
                                iters_ptr=None, fevals_ptr=None, status_ptr=None, ntrace=0,
                                trace_ptr=None, flags=0, stream=None):
        ls = ls or BackTracking()
        lsc = ls._c()
        h = self.model_handle(obj)
        _lib.check(self._lib.qn_bfgs_optimize_batched_device_f64(
            self._ctx, h, n, batch, theta_ptr, theta_stride or n, C.byref(lsc), tol, max_iter,
            states_ptr, state_stride, ld, x_ptr, g_ptr, ell_ptr, iters_ptr, fevals_ptr, status_ptr,
            ntrace, trace_ptr, flags, stream))

    def logdensity_batched(self, obj, Theta, grad=False):
        Theta = np.asfortranarray(Theta, dtype=np.float64)
        n, batch = Theta.shape
        h = self.model_handle(obj)
        ell = np.empty(batch)
        g = np.empty((n, batch), order="F") if grad else None
        _lib.check(self._lib.qn_logdensity_batched_f64(self._ctx, h, n, batch, _ptr(Theta), _ptr(ell), _ptr(g)))
        return (ell, g) if grad else ell


def optimize_batched_multi(engines, obj, Theta, ls=None, tol=1e-8, max_iter=MAX_ITER_DEFAULT, flags=0,
                           out=None):
    """One batch sharded over several GPUs from ONE process (qn_bfgs_optimize_batched_multi_f64;
    SURVEY §8(e) row 1): problem j goes to engines[j mod G], no inter-GPU communication, the host
    scatters the starts and gathers the results.  `out`: optional dict of preallocated (pinned)
    arrays x, g (n×batch, column-major), ell, iters, fevals, status to write into."""
    Theta = np.asfortranarray(Theta, dtype=np.float64)
    n, batch = Theta.shape
    ls = ls or BackTracking()
    lsc = ls._c()
    lib = engines[0]._lib
    G = len(engines)
    ctxs = (C.c_void_p * G)(*[e._ctx for e in engines])
    models = (C.c_void_p * G)(*[e.model_handle(obj) for e in engines])
    out = out or {}
    x = out.get("x") if out.get("x") is not None else np.empty((n, batch), order="F")
    g = out.get("g") if out.get("g") is not None else np.empty((n, batch), order="F")
    ell = out.get("ell") if out.get("ell") is not None else np.empty(batch)
    iters = out.get("iters") if out.get("iters") is not None else np.empty(batch, dtype=np.int32)
    fevals = out.get("fevals") if out.get("fevals") is not None else np.empty(batch, dtype=np.int32)
    status = out.get("status") if out.get("status") is not None else np.empty(batch, dtype=np.int32)
    _lib.check(lib.qn_bfgs_optimize_batched_multi_f64(
        ctxs, G, models, n, batch, _ptr(Theta), C.byref(lsc), tol, max_iter, _ptr(x), _ptr(g), _ptr(ell),
        _ptr(iters), _ptr(fevals), _ptr(status), flags))
    return dict(x=x, g=g, ell=ell, iters=iters, fevals=fevals, status=status)


def optimize_rowsharded(engines, obj, theta, ls=None, tol=1e-8, max_iter=MAX_ITER_DEFAULT, ntrace=0,
                        flags=0, want_h=False, bench_iters=0):
    """ONE large problem on several GPUs: engine g holds the rows of its row tiles of B⁻¹
    (qn_bfgs_optimize_rowsharded_f64).  Same result as one GPU, bit for bit.
    Returns a dict like Engine.optimize_batched with batch = 1 (plus 'H', 'ms_per_iter')."""
    theta = np.ascontiguousarray(theta, dtype=np.float64).reshape(-1)
    n = theta.size
    ls = ls or BackTracking()
    lsc = ls._c()
    lib = engines[0]._lib
    h = engines[0].model_handle(obj)
    ctxs = (C.c_void_p * len(engines))(*[e._ctx for e in engines])
    x = np.empty(n); g = np.empty(n)
    H = np.zeros((n, n), order="F") if want_h else None
    ell = C.c_double(float("nan")); it = C.c_int32(0); fe = C.c_int32(0); st = C.c_int32(0)
    ms = C.c_double(0.0)
    trace = np.full((ntrace, n), np.nan) if ntrace else None
    _lib.check(lib.qn_bfgs_optimize_rowsharded_f64(
        ctxs, len(engines), h, n, _ptr(theta), C.byref(lsc), tol, max_iter, _ptr(x), _ptr(g), _ptr(H), n,
        C.byref(ell), C.byref(it), C.byref(fe), C.byref(st), ntrace, _ptr(trace), flags, bench_iters,
        C.byref(ms)))
    return dict(x=x[:, None], g=g[:, None], ell=np.array([ell.value]), iters=np.array([it.value], dtype=np.int32),
                fevals=np.array([fe.value], dtype=np.int32), status=np.array([st.value], dtype=np.int32),
                trace=None if trace is None else trace[None], H=H, ms_per_iter=ms.value)


_engines = {}
_engines_lock = threading.Lock()


def default_engine(device=0):
    with _engines_lock:
        if device not in _engines:
            _engines[device] = Engine(device)
        return _engines[device]


# ------------------------------------------------------------------ the reference's functions
def optimize_(state, obj, x, ls=None, tol=1e-8, engine=None):
    """optimize!(state, obj, x, ls=BackTracking(), tol=1e-8) (QN.jl:237).

    Single problem: `state` a BFGSState, `x` a length-P vector (not modified,
    QN.jl:248); returns ℓ at the optimum or NaN (QN.jl:261,291).
    Batched: `state` a BFGSStateBatch, `x` a P×batch matrix (one start per
    column); returns the vector of ℓ / NaN.
    """
    eng = engine or default_engine()
    ls = ls or BackTracking()
    if isinstance(state, BFGSStateBatch):
        r = eng.optimize_batched(obj, x, ls=ls, tol=tol, states=state, want_xg=False)
        return r["ell"]
    x = np.ascontiguousarray(x, dtype=np.float64)
    if x.size != state.P:
        raise ValueError("x has length %d, state is BFGSState{%d}" % (x.size, state.P))
    h = eng.model_handle(obj)
    lsc = ls._c()
    ell = C.c_double()
    it, fe, st = C.c_int32(), C.c_int32(), C.c_int32()
    _lib.check(eng._lib.qn_bfgs_optimize_f64(eng._ctx, h, state.P, _ptr(x), C.byref(lsc), tol,
                                             _ptr(state.data), state.L, C.byref(ell), C.byref(it),
                                             C.byref(fe), C.byref(st), 0))
    state.iters, state.fevals, state.status = it.value, fe.value, st.value
    return ell.value


optimize = optimize_


def logdensity(obj, theta, engine=None):
    """logdensity(obj, θ) (QN.jl:8), evaluated by the registered device objective."""
    eng = engine or default_engine()
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    return float(eng.logdensity_batched(obj, theta.reshape(-1, 1))[0])


def dlogdensity_(grad, obj, theta, engine=None):
    """∂logdensity!(∇, obj, θ) (QN.jl:9): fills `grad` in place, returns ℓ."""
    eng = engine or default_engine()
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    ell, g = eng.logdensity_batched(obj, theta.reshape(-1, 1), grad=True)
    grad[...] = g[:, 0]
    return float(ell[0])

"""ctypes binding for the CPU oracle (oracle/libqn_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package never
imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libqn_oracle.so")

ROSENBROCK, DENSE_GAUSSIAN, LOGISTIC = 0, 1, 2
CONVERGED, NONFINITE, LINESEARCH_FAILED, MAX_ITER = 0, 1, 2, 3
RBLOCK = 0  # canonical rule C2 blocks (qo_chain_block)
ONEPASS = -1  # rblock value selecting the one-pass deferred-update variant (not the reference order)


class Backtracking(C.Structure):
    _fields_ = [("c1", C.c_double), ("rho_hi", C.c_double), ("rho_lo", C.c_double),
                ("iterations", C.c_int64), ("order", C.c_int32), ("_pad", C.c_int32)]


class Gaussian(C.Structure):
    _fields_ = [("mu", C.c_void_p), ("lam", C.c_void_p)]


class Logistic(C.Structure):
    _fields_ = [("m", C.c_int32), ("_pad", C.c_int32), ("X", C.c_void_p), ("Xt", C.c_void_p),
                ("y", C.c_void_p), ("lam", C.c_double)]


class Model(C.Structure):
    _fields_ = [("kind", C.c_int32), ("_pad", C.c_int32), ("params", C.c_void_p)]


class Info(C.Structure):
    _fields_ = [("iters", C.c_int64), ("fevals", C.c_int64), ("gevals", C.c_int64),
                ("status", C.c_int32), ("_pad", C.c_int32)]


_FAST_SO = os.path.join(_HERE, "libqn_fastport.so")


def _stale(so, deps):
    return not os.path.exists(so) or any(
        os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(so) for f in deps)


def build(force=False):
    if force or _stale(_SO, ("qn_oracle.c", "qn_oracle.h", "Makefile")):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libqn_oracle.so"])
    if force or _stale(_FAST_SO, ("qn_fastport.c", "qn_oracle.h", "Makefile")):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libqn_fastport.so"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        dp = C.POINTER(C.c_double)
        L.qo_state_ld.restype = C.c_int
        L.qo_state_ld.argtypes = [C.c_int, C.c_int]
        L.qo_state_len.restype = C.c_int64
        L.qo_state_len.argtypes = [C.c_int, C.c_int]
        L.qo_logdensity.restype = C.c_double
        L.qo_logdensity.argtypes = [C.POINTER(Model), C.c_int, dp]
        L.qo_dlogdensity.restype = C.c_double
        L.qo_dlogdensity.argtypes = [dp, C.POINTER(Model), C.c_int, dp]
        L.qo_bfgs_update.restype = C.c_double
        L.qo_bfgs_update.argtypes = [C.c_int, C.c_int, dp, dp, dp, dp, dp, dp, C.c_int, dp]
        L.qo_optimize.restype = C.c_double
        L.qo_optimize.argtypes = [C.c_int, C.c_int, dp, C.POINTER(Model), dp,
                                  C.POINTER(Backtracking), C.c_double, C.c_int64, C.c_int,
                                  C.c_int, dp, dp, C.POINTER(Info)]
        L.qo_optimize_batch.restype = C.c_int64
        L.qo_optimize_batch.argtypes = [C.c_int, C.c_int64, dp, C.POINTER(Model),
                                        C.POINTER(Backtracking), C.c_double, C.c_int64, C.c_int,
                                        dp, dp, dp, C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                        C.POINTER(C.c_int32), C.c_int, dp, C.c_int]
        L.qo_num_cores.restype = C.c_int
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def backtracking(order=2, c1=1e-4, rho_hi=0.5, rho_lo=0.1, iterations=1000):
    return Backtracking(c1, rho_hi, rho_lo, iterations, order, 0)


class OracleModel:
    """Keeps the numpy arrays a qo_model points into alive."""

    def __init__(self, kind, **kw):
        self.kind = kind
        self._keep = []
        self._params = None
        if kind == DENSE_GAUSSIAN:
            mu = np.ascontiguousarray(kw["mu"], dtype=np.float64)
            lam = np.asfortranarray(kw["lam"], dtype=np.float64)
            self._keep += [mu, lam]
            self._params = Gaussian(mu.ctypes.data, lam.ctypes.data)
        elif kind == LOGISTIC:
            X = np.asfortranarray(kw["X"], dtype=np.float64)          # m x n, ld = m
            Xt = np.asfortranarray(X.T.copy(order="F"))               # n x m, ld = n
            y = np.ascontiguousarray(kw["y"], dtype=np.float64)
            self._keep += [X, Xt, y]
            self._params = Logistic(X.shape[0], 0, X.ctypes.data, Xt.ctypes.data, y.ctypes.data,
                                    float(kw["lam"]))
        self.c = Model(kind, 0, C.addressof(self._params) if self._params is not None else None)


def rosenbrock():
    return OracleModel(ROSENBROCK)


def state_ld(n, native_w=8):
    return lib().qo_state_ld(n, native_w)


def logdensity(model, theta):
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    return lib().qo_logdensity(C.byref(model.c), theta.size, _dp(theta))


def dlogdensity(model, theta):
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    g = np.empty_like(theta)
    v = lib().qo_dlogdensity(_dp(g), C.byref(model.c), theta.size, _dp(theta))
    return v, g


def bfgs_update(H, s, g_new, g_old, rblock=RBLOCK):
    """H: (n x n) array interpreted as H[c, r]; returns (m, H_new, u, d)."""
    n = s.size
    Hf = np.asfortranarray(H, dtype=np.float64).copy(order="F")
    u = np.zeros(n)
    y = np.zeros(n)
    scratch = np.zeros(5 * n + 64)
    m = lib().qo_bfgs_update(n, n, _dp(Hf), _dp(u), _dp(y), _dp(np.ascontiguousarray(s)),
                             _dp(np.ascontiguousarray(g_new)), _dp(np.ascontiguousarray(g_old)),
                             rblock, _dp(scratch))
    return m, Hf, u, y


def optimize(model, x0, ls=None, tol=1e-8, max_iter=10000, L=None, ntrace=0, rblock=RBLOCK):
    """Single problem.  Returns dict(ell, x, g, state, iters, fevals, status, trace, trace_scalars)."""
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    n = x0.size
    L = L or state_ld(n)
    ls = ls or backtracking()
    state = np.full(lib().qo_state_len(n, L), np.nan)
    trace = np.full((ntrace, n), np.nan) if ntrace else None
    tsc = np.full((ntrace, 4), np.nan) if ntrace else None
    info = Info()
    ell = lib().qo_optimize(n, L, _dp(state), C.byref(model.c), _dp(x0), C.byref(ls), tol, max_iter,
                            rblock, ntrace, _dp(trace), _dp(tsc), C.byref(info))
    return dict(ell=ell, x=state[:n].copy(), g=state[L:L + n].copy(), state=state, L=L,
                iters=info.iters, fevals=info.fevals, gevals=info.gevals, status=info.status,
                trace=trace, trace_scalars=tsc)


def optimize_batch(model, theta0, ls=None, tol=1e-8, max_iter=10000, ntrace=0, nthreads=0,
                   rblock=RBLOCK, want_xg=True):
    """theta0: (n x batch) column-major (one problem per column)."""
    theta0 = np.asfortranarray(theta0, dtype=np.float64)
    n, batch = theta0.shape
    ls = ls or backtracking()
    x = np.empty((n, batch), order="F") if want_xg else None
    g = np.empty((n, batch), order="F") if want_xg else None
    ell = np.empty(batch)
    iters = np.empty(batch, dtype=np.int64)
    fevals = np.empty(batch, dtype=np.int64)
    status = np.empty(batch, dtype=np.int32)
    trace = np.full((batch, ntrace, n), np.nan) if ntrace else None
    nconv = lib().qo_optimize_batch(
        n, batch, _dp(theta0), C.byref(model.c), C.byref(ls), tol, max_iter, rblock, _dp(x), _dp(g),
        _dp(ell), iters.ctypes.data_as(C.POINTER(C.c_int64)),
        fevals.ctypes.data_as(C.POINTER(C.c_int64)), status.ctypes.data_as(C.POINTER(C.c_int32)),
        ntrace, _dp(trace), nthreads)
    return dict(x=x, g=g, ell=ell, iters=iters, fevals=fevals, status=status, trace=trace,
                converged=int(nconv))


def num_cores():
    return lib().qo_num_cores()


# ---------------------------------------------------------------------------- the fast port
# oracle/qn_fastport.c: the same algorithm with the freedoms @turbo / @fastmath take (free summation
# order, contraction).  CPU BASELINE ONLY — never the parity checker (tests hold it to the
# north-star tolerances against the strict oracle above).
_fast = None


def fast_lib():
    global _fast
    if _fast is None:
        if not os.path.exists(_FAST_SO):
            build()
        L = C.CDLL(_FAST_SO)
        dp = C.POINTER(C.c_double)
        L.qf_optimize.restype = C.c_double
        L.qf_optimize.argtypes = [C.c_int, dp, C.POINTER(Backtracking), C.c_double, C.c_int64, dp, dp,
                                  C.c_int, dp, C.POINTER(Info)]
        L.qf_time_single.restype = None
        L.qf_time_single.argtypes = [C.c_int, dp, C.POINTER(Backtracking), C.c_double, C.c_int, dp]
        L.qf_optimize_batch.restype = C.c_int64
        L.qf_optimize_batch.argtypes = [C.c_int, C.c_int64, dp, C.POINTER(Backtracking), C.c_double,
                                        C.c_int64, dp, dp, dp, C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                        C.POINTER(C.c_int32), C.c_int, dp, C.c_int]
        _fast = L
    return _fast


def fast_optimize_batch(theta0, ls=None, tol=1e-8, max_iter=10000, ntrace=0, nthreads=0, want_xg=True):
    """Rosenbrock batch through the fast port; same result dict as optimize_batch."""
    theta0 = np.asfortranarray(theta0, dtype=np.float64)
    n, batch = theta0.shape
    ls = ls or backtracking()
    x = np.empty((n, batch), order="F") if want_xg else None
    g = np.empty((n, batch), order="F") if want_xg else None
    ell = np.empty(batch)
    iters = np.empty(batch, dtype=np.int64)
    fevals = np.empty(batch, dtype=np.int64)
    status = np.empty(batch, dtype=np.int32)
    trace = np.full((batch, ntrace, n), np.nan) if ntrace else None
    nconv = fast_lib().qf_optimize_batch(
        n, batch, _dp(theta0), C.byref(ls), tol, max_iter, _dp(x), _dp(g), _dp(ell),
        iters.ctypes.data_as(C.POINTER(C.c_int64)), fevals.ctypes.data_as(C.POINTER(C.c_int64)),
        status.ctypes.data_as(C.POINTER(C.c_int32)), ntrace, _dp(trace), nthreads)
    return dict(x=x, g=g, ell=ell, iters=iters, fevals=fevals, status=status, trace=trace, converged=int(nconv))


def fast_time_single(x0, reps=200, ls=None, tol=1e-8):
    """Per-call seconds of `reps` solves of one problem, timed inside C (BenchmarkTools-style)."""
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    ls = ls or backtracking()
    sec = np.empty(reps)
    fast_lib().qf_time_single(x0.size, _dp(x0), C.byref(ls), tol, reps, _dp(sec))
    return sec

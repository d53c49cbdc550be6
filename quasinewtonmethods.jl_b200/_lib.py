"""ctypes binding of libqn_cuda.so (include/qn_cuda.h).

The shared library is the product; this module only declares its entry points.
There is no fallback of any kind: if the library is missing, or no B200 is
visible, the calls raise.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# QN_LIBRARY_PATH: a library built with a plug-in objective (include/qn_objective.cuh)
SO_PATH = os.environ.get("QN_LIBRARY_PATH") or os.path.join(_HERE, "libqn_cuda.so")

# error codes (include/qn_cuda.h)
OK, ERR_INVALID_ARGUMENT, ERR_CUDA, ERR_UNREGISTERED_OBJECTIVE, ERR_UNSUPPORTED_SIZE, ERR_NO_DEVICE = range(6)
# per-problem status
CONVERGED, NONFINITE, LINESEARCH_FAILED, MAX_ITER = range(4)
# objective kinds
OBJ_ROSENBROCK, OBJ_DENSE_GAUSSIAN, OBJ_LOGISTIC = range(3)
OBJ_USER = 100
# flags
FLAG_WRITE_H, FLAG_FORCE_SMEM, FLAG_FORCE_STREAM, FLAG_ONE_PASS = 0x1, 0x2, 0x4, 0x8
# execution paths reported by qn_last_launch_info
PATH_REGISTER, PATH_SMEM, PATH_STREAM = 0, 1, 2


class QNError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libqn_cuda error %d: %s" % (code, msg))
        self.code = code


class UnregisteredObjectiveError(QNError):
    """The MethodError analogue (QN.jl:254): the objective has no device implementation."""


class Backtracking(C.Structure):
    _fields_ = [("c1", C.c_double), ("rho_hi", C.c_double), ("rho_lo", C.c_double),
                ("iterations", C.c_int64), ("order", C.c_int32), ("_pad", C.c_int32)]


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_vp = C.c_void_p

# name -> (restype, argtypes); the export test checks every name in qn_cuda.h is here
SIGNATURES = {
    "qn_version": (C.c_int32, []),
    "qn_last_error": (C.c_char_p, []),
    "qn_device_count": (C.c_int32, []),
    "qn_backtracking_default": (None, [C.POINTER(Backtracking), C.c_int32]),
    "qn_create": (C.c_int32, [C.POINTER(_vp), C.c_int32]),
    "qn_destroy": (C.c_int32, [_vp]),
    "qn_device_info": (C.c_int32, [_vp, C.c_char_p, C.c_size_t, _ip, _ip, _ip]),
    "qn_set_option": (C.c_int32, [_vp, C.c_char_p, C.c_int64]),
    "qn_get_option": (C.c_int32, [_vp, C.c_char_p, C.POINTER(C.c_int64)]),
    "qn_last_tail_profile": (C.c_int32, [_vp, _dp, _dp, _ip]),
    "qn_host_alloc": (C.c_int32, [C.POINTER(_vp), C.c_size_t]),
    "qn_host_free": (C.c_int32, [_vp]),
    "qn_state_ld": (C.c_int32, [C.c_int32, C.c_int32]),
    "qn_state_len": (C.c_int64, [C.c_int32, C.c_int32]),
    "qn_model_create": (C.c_int32, [_vp, C.c_int32, C.POINTER(_vp)]),
    "qn_model_dense_gaussian": (C.c_int32, [_vp, C.c_int32, _vp, _vp, C.POINTER(_vp)]),
    "qn_model_logistic": (C.c_int32, [_vp, C.c_int32, C.c_int32, _vp, _vp, C.c_double, C.POINTER(_vp)]),
    "qn_model_user": (C.c_int32, [_vp, C.c_int32, _vp, C.c_size_t, C.POINTER(_vp)]),
    "qn_model_free": (C.c_int32, [_vp]),
    "qn_logdensity_batched_f64": (C.c_int32, [_vp, _vp, C.c_int32, C.c_int64, _vp, _vp, _vp]),
    "qn_bfgs_optimize_batched_f64": (C.c_int32, [
        _vp, _vp, C.c_int32, C.c_int64, _vp, C.POINTER(Backtracking), C.c_double, C.c_int32,
        _vp, C.c_int64, C.c_int32, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int32, _vp, C.c_uint32]),
    "qn_bfgs_optimize_batched_device_f64": (C.c_int32, [
        _vp, _vp, C.c_int32, C.c_int64, _vp, C.c_int64, C.POINTER(Backtracking), C.c_double,
        C.c_int32, _vp, C.c_int64, C.c_int32, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int32, _vp,
        C.c_uint32, _vp]),
    "qn_bfgs_optimize_f64": (C.c_int32, [
        _vp, _vp, C.c_int32, _vp, C.POINTER(Backtracking), C.c_double, _vp, C.c_int32, _dp, _ip,
        _ip, _ip, C.c_uint32]),
    "qn_bfgs_optimize_batched_multi_f64": (C.c_int32, [
        C.POINTER(_vp), C.c_int32, C.POINTER(_vp), C.c_int32, C.c_int64, _vp,
        C.POINTER(Backtracking), C.c_double, C.c_int32, _vp, _vp, _vp, _vp, _vp, _vp, C.c_uint32]),
    "qn_bfgs_optimize_rowsharded_f64": (C.c_int32, [
        C.POINTER(_vp), C.c_int32, _vp, C.c_int32, _vp, C.POINTER(Backtracking), C.c_double, C.c_int32,
        _vp, _vp, _vp, C.c_int32, _dp, _ip, _ip, _ip, C.c_int32, _vp, C.c_uint32, C.c_int32, _dp]),
    "qn_last_launch_info": (C.c_int32, [_vp, _ip, _ip, _ip, _ip, _ip]),
    "qn_microbench_dfma": (C.c_int32, [_vp, _dp]),
    "qn_microbench_lds": (C.c_int32, [_vp, _dp]),
    "qn_microbench_hbm": (C.c_int32, [_vp, _dp]),
    "qn_stream_bench_f64": (C.c_int32, [_vp, C.c_int32, C.c_int32, C.c_int32, C.c_uint32, _dp, _dp]),
}

_lib = None


def load():
    """dlopen the in-tree libqn_cuda.so; raise if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise ImportError(
                "%s not found: build it with `python __graft_entry__.py` (or "
                "quasinewtonmethods.jl_b200/build.py); there is no CPU fallback" % SO_PATH)
        lib = C.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            if not hasattr(lib, name):
                continue  # the export test reports what is missing
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc):
    if rc != OK:
        msg = load().qn_last_error().decode("utf-8", "replace")
        if rc == ERR_UNREGISTERED_OBJECTIVE:
            raise UnregisteredObjectiveError(rc, msg)
        raise QNError(rc, msg)

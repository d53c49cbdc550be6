"""Build recipe for libqn_cuda.so (sm_100a only, in-tree, no torch dependency).

    python "quasinewtonmethods.jl_b200/build.py" [--force] [--verbose]

-fmad=false is REQUIRED: the canonical arithmetic order fixes
where fused multiply-adds happen; nvcc must never contract on its own.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
SO = os.path.join(HERE, "libqn_cuda.so")

SOURCES = ["qn_capi.cu", "qn_launch_tiles.cu", "qn_launch_small.cu", "qn_eval.cu", "qn_multi.cu", "qn_microbench.cu", "qn_stream.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo", "-fmad=false",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O2",
    "--expt-relaxed-constexpr",
] + os.environ.get("QN_EXTRA_NVCC", "").split()


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _sources():
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def _stale():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(ROOT, "include", "qn_cuda.h"))
    deps.append(os.path.abspath(__file__))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile every CUDA source into quasinewtonmethods.jl_b200/libqn_cuda.so.

    QN_USER_OBJECTIVE=/abs/path/my.cuh builds a library WITH that plug-in objective
    (include/qn_objective.cuh) into QN_SO_OUT (default: libqn_cuda_user.so next to the stock
    library); the stock libqn_cuda.so is left alone."""
    user = os.environ.get("QN_USER_OBJECTIVE")
    if user:
        return _build(os.environ.get("QN_SO_OUT", os.path.join(HERE, "libqn_cuda_user.so")), "build_user",
                      ["-DQN_USER_OBJECTIVE_HEADER=\"%s\"" % os.path.abspath(user)], verbose)
    if not force and not _stale():
        return SO
    return _build(SO, "build", [], verbose)


def _build(so, subdir, extra, verbose):
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, subdir), exist_ok=True)
    for src in _sources():
        obj = os.path.join(HERE, subdir, os.path.basename(src) + ".o")
        cmd = [_nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out = p.communicate()[0].decode()
        if p.returncode != 0 or verbose:
            sys.stderr.write("== %s\n%s\n" % (os.path.basename(src), out))
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    subprocess.check_call([_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", so] + objs + ["-lcudart_static", "-lpthread", "-ldl", "-lrt"])
    return so


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))

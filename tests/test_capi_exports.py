"""libqn_cuda.so loads without a GPU and exports every symbol include/qn_cuda.h declares."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def declared_functions():
    src = open(os.path.join(ROOT, "include", "qn_cuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(qn_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_path():
    names = declared_functions()
    for must in ("qn_create", "qn_bfgs_optimize_f64", "qn_bfgs_optimize_batched_f64",
                 "qn_bfgs_optimize_batched_device_f64", "qn_model_create", "qn_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol(qn):
    lib = ctypes.CDLL(qn._lib.SO_PATH)
    missing = [n for n in declared_functions() if not hasattr(lib, n)]
    assert not missing, missing


def test_binding_covers_every_declared_symbol(qn):
    missing = [n for n in declared_functions() if n not in qn._lib.SIGNATURES]
    assert not missing, missing


def test_no_compute_without_gpu_but_metadata_works(qn):
    lib = qn._lib.load()
    assert lib.qn_version() == 100
    # BFGSState layout helpers are pure host functions (QN.jl:84-104)
    assert lib.qn_state_ld(60, 8) == 64 and lib.qn_state_ld(60, 4) == 60
    assert lib.qn_state_len(60, 64) * 8 == 34304
    ls = qn._lib.Backtracking()
    lib.qn_backtracking_default(ctypes.byref(ls), 3)
    assert (ls.c1, ls.rho_hi, ls.rho_lo, ls.iterations, ls.order) == (1e-4, 0.5, 0.1, 1000, 3)


def test_product_has_no_cpu_fallback(qn):
    """Without a CUDA device the product path fails loudly instead of computing on the CPU."""
    lib = qn._lib.load()
    if lib.qn_device_count() > 0:
        pytest.skip("a GPU is visible; the no-device path is exercised on the CPU box")
    with pytest.raises(qn.QNError) as e:
        qn.Engine(0)
    assert e.value.code == qn._lib.ERR_NO_DEVICE
    assert "no CPU fallback" in str(e.value)


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under the package may reference it."""
    pkg = os.path.join(ROOT, "quasinewtonmethods.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        if os.path.basename(dirpath) in ("build", "__pycache__"):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".inc", ".jl")):
                text = open(os.path.join(dirpath, f), encoding="utf-8").read()
                assert "qn_oracle" not in text and "from oracle" not in text and "import oracle" not in text, f


def test_headline_kernel_uses_tensor_memory(qn):
    """The sm_100a SASS of the n = 60 kernel that large batches run moves its B⁻¹ tiles with tensor-memory
    loads and stores (LDTM / STTM — what tcgen05.ld / tcgen05.st compile to) and computes in fp64 FMAs."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    name = "_ZN2qn28qn_resident_tile_tmem_kernelILi60ELi1EEEvNS_11BatchParamsE"
    out = subprocess.run([cuobjdump, "-sass", "-fun", name, qn._lib.SO_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out or "sm_100" in out
    assert out.count("LDTM") >= 30 and out.count("STTM") >= 15, (out.count("LDTM"), out.count("STTM"))
    assert out.count("DFMA") >= 300
    assert "UTCATOMSWS" in out  # tcgen05.alloc

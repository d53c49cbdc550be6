"""The device-objective extension point (include/qn_objective.cuh): build a
library WITH examples/quartic_objective.cuh, then solve with it through the ordinary entry points.
The stock library must refuse the same model (no CPU fallback)."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from conftest import ROOT, starts

pytestmark = pytest.mark.gpu


def test_stock_library_refuses_user_objective(eng, qn):
    with pytest.raises(qn.UnregisteredObjectiveError):
        eng.optimize_batched(qn.UserObjective(np.zeros(3 * 5), n=5), starts(5, 2))


def test_plugin_objective_builds_and_solves(tmp_path):
    so = str(tmp_path / "libqn_cuda_quartic.so")
    env = dict(os.environ, QN_USER_OBJECTIVE=os.path.join(ROOT, "examples", "quartic_objective.cuh"), QN_SO_OUT=so)
    subprocess.check_call([sys.executable, os.path.join(ROOT, "quasinewtonmethods.jl_b200", "build.py")], env=env)
    assert os.path.exists(so)
    code = textwrap.dedent("""
        import sys, numpy as np
        sys.path.insert(0, %r); sys.path.insert(0, %r)
        import qn_b200 as qn
        from conftest import starts
        for n in (7, 40, 128):
            rng = np.random.Generator(np.random.Philox(key=n))
            a = rng.uniform(0.1, 1.0, n); b = rng.uniform(0.5, 2.0, n); c = rng.standard_normal(n)
            obj = qn.UserObjective(np.concatenate([a, b, c]), n=n)
            eng = qn.default_engine(0)
            Th = starts(n, 9, key=n)
            ell, g = eng.logdensity_batched(obj, Th, grad=True)
            d = Th - c[:, None]
            assert np.allclose(ell, -(a[:, None] * d**4 + 0.5 * b[:, None] * d**2).sum(0), rtol=1e-12)
            assert np.allclose(g, -(4 * a[:, None] * d**3 + b[:, None] * d), rtol=1e-12, atol=1e-13)
            r = eng.optimize_batched(obj, Th)
            assert (r["status"] == qn.CONVERGED).all(), r["status"]
            assert np.abs(r["g"]).max() < 1e-8
            assert np.abs(r["x"] - c[:, None]).max() < 1e-6
            print("quartic n=%%d: iterations %%s" %% (n, r["iters"]))
        print("plugin ok")
    """ % (ROOT, os.path.join(ROOT, "tests")))
    out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, QN_LIBRARY_PATH=so),
                         capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "plugin ok" in out.stdout

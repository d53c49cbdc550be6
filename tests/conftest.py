import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SEED = 20261018
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def starts(n, batch, key=SEED):
    """n×batch column-major N(0,1) starts from the counter-based Philox generator."""
    rng = np.random.Generator(np.random.Philox(key=key))
    return rng.standard_normal((batch, n)).T.copy(order="F")


@pytest.fixture(scope="session")
def ob():
    """The CPU oracle (the checker)."""
    from oracle import binding
    binding.build()
    return binding


@pytest.fixture(scope="session")
def qn():
    import qn_b200
    return qn_b200


@pytest.fixture(scope="session")
def eng(qn):
    """The product engine on cuda:0.  Fails loudly when the CUDA path is unavailable."""
    return qn.default_engine(0)


@pytest.fixture
def option(eng):
    """Set options of the session's context for one test (qn_set_option) and restore them after."""
    saved = {}

    def set_(key, value):
        if key not in saved:
            saved[key] = eng.get_option(key)
        eng.set_option(key, value)

    yield set_
    for k, v in saved.items():
        eng.set_option(k, v)

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import __graft_entry__ as graft  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "ref: needs oracle/_ref (the reference built behind shims)")


@pytest.fixture(scope="session")
def pkg():
    return graft.load_package()


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the oracle library and liblpgpu.so exist (they are build products, not in git)."""
    lib = os.path.join(graft.PKG_DIR, "liblpgpu.so")
    ora = os.path.join(ROOT, "oracle", "liblp_oracle.so")
    if not (os.path.exists(lib) and os.path.exists(ora)):
        graft.build()


def relerr(a, b):
    import numpy as np
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(float(np.max(np.abs(b))), 1e-300))

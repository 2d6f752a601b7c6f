import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "mpc-mmd_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(1, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built():
    """in-tree native artefacts (libmpcmmd.so, liboracle.so); nvcc/gcc only run when sources changed"""
    import __graft_entry__ as g
    g.build()
    return g

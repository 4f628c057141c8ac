import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def pkg():
    """The product package (host mirror + libmgn_b200.so).  Built on demand, never mocked."""
    lib = os.path.join(ROOT, "meshgraphnets.jl_b200", "csrc", "libmgn_b200.so")
    if not os.path.exists(lib):
        import __graft_entry__
        __graft_entry__.build()
    import mgn_pkg
    return mgn_pkg.pkg

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "multigpu: needs at least two CUDA devices")


def _cuda_devices():
    try:
        import torch
        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    """A host without a CUDA device SKIPS the `gpu` tests (plain `pytest -q` stays green on CPU-only CI); a GPU host
    runs them and they fail loudly if the library is missing - there is no fallback to skip to."""
    n = _cuda_devices()
    no_gpu = pytest.mark.skip(reason="needs a CUDA device")
    one_gpu = pytest.mark.skip(reason="needs at least two CUDA devices")
    for item in items:
        if "gpu" in item.keywords and n == 0:
            item.add_marker(no_gpu)
        elif "multigpu" in item.keywords and n < 2:
            item.add_marker(one_gpu)


@pytest.fixture(scope="session")
def pkg():
    """The product package (host mirror + libmgn_b200.so).  Built on demand, never mocked."""
    lib = os.path.join(ROOT, "meshgraphnets.jl_b200", "csrc", "libmgn_b200.so")
    if not os.path.exists(lib):
        import __graft_entry__
        __graft_entry__.build()
    import mgn_pkg
    return mgn_pkg.pkg

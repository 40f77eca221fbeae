import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _fresh_library():
    """The tests load ``xlxmert_b200/lib/libxlxmert_b200.so`` as it lies in the tree; rebuild it first whenever the CUDA
    sources changed since it was built (a no-op when the build stamp matches), so that a stale binary can never be what
    a green or red result is about."""
    import __graft_entry__ as entry
    entry.build()

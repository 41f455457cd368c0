import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """The tests call through libqmcb.so; build it in-tree if a fresh checkout has none."""
    from qmctorch_b200 import build as b
    if not os.path.isfile(b.LIB):
        b.build()
    yield


@pytest.fixture
def double_default():
    """torch default dtype float64 for the duration of a test (the reference runs under
    set_torch_double_precision(); CPU-generator draws depend on the default dtype)."""
    import torch
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    yield
    torch.set_default_dtype(old)

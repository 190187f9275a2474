import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (test infrastructure, oracle/)."""
    import oracle
    return oracle


@pytest.fixture(scope="session")
def gb():
    """The product package, loaded through the C ABI; requires a GPU."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import grace_devel_b200
    return grace_devel_b200

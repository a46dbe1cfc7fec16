import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def ctx():
    import torch
    import gridaphybrid_b200 as gh
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    c = gh.Context(0)
    gh.set_default_context(c)
    yield c

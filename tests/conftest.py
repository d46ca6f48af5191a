import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "kat.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def orc():
    import oracle
    oracle.port()   # builds oracle/_build on first use (gcc only)
    return oracle


@pytest.fixture(scope="session")
def jp():
    """The product: libjpbwt.so through its Python mirror (GPU tests only; there is no CPU path to fall back on)."""
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import jampack_b200
    from jampack_b200 import build
    build.build()
    assert jampack_b200.device_count() >= 1
    return jampack_b200

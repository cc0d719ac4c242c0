import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

MODELS = os.path.join(ROOT, "tests", "golden", "models")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def oscene():
    from oracle import scene as S
    return S


@pytest.fixture(scope="session")
def device():
    """A pbr_ctx on cuda:0.  Fails loudly (no skip, no fallback) when the library or device is missing."""
    import pbr_b200
    d = pbr_b200.Device(0)
    yield d
    d.close()

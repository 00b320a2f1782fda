import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import Oracle, build

    build()
    return Oracle()


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    return np.load(ROOT / "tests" / "golden" / "flann_nn.npz")

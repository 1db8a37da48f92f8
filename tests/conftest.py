import pathlib
import sys

import pytest

ROOT = pathlib.Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    return np.load(ROOT / "tests" / "golden" / "sampling_ref.npz")


@pytest.fixture(scope="session")
def twin():
    from oracle import twin as t

    t.lib()
    return t

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run through gpurun)")


@pytest.fixture(scope="session")
def golden_small():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_small.npz"))


@pytest.fixture(scope="session")
def golden_hashes():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "golden_hashes.json")) as f:
        return json.load(f)

import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    path = os.path.join(ROOT, "tests", "golden", "reference_outputs.npz")
    with np.load(path) as data:
        return {k: data[k] for k in data.files}


def has_gpu():
    try:
        from structurepreservingiterativesolvers_b200 import _native as nat
        return nat.device_count() > 0
    except Exception:
        return False

import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def lrh_golden_cases():
    z = load_golden("lrh.npz")
    idx = sorted({k.split("/")[0] for k in z.files})
    cases = []
    for i in idx:
        cases.append({k.split("/")[1]: z[k] for k in z.files if k.startswith(i + "/")})
    return cases


@pytest.fixture(scope="session")
def lrh_cases():
    return lrh_golden_cases()

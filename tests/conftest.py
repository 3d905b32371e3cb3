import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


def golden_cases(npz):
    """Groups 'case__field' keys of an npz into {case: {field: array}}."""
    out = {}
    for key in npz.files:
        case, field = key.split("__", 1)
        out.setdefault(case, {})[field] = npz[key]
    return out


@pytest.fixture(scope="session")
def pcrv_cases():
    return golden_cases(load_golden("pcrv_cases"))


@pytest.fixture(scope="session")
def fit_cases():
    return golden_cases(load_golden("fits"))


def case_mis_cfs(case):
    n = int(case["npdim"])
    return [case[f"mi{j}"] for j in range(n)], [case[f"cf{j}"] for j in range(n)]


def case_types(case):
    return [str(t) for t in case["types"]]

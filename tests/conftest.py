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


def c1_inputs(n=100_000, d=10):
    """BASELINE.json configs[0] / SURVEY 8(d) C1 synthetic inputs, as tests/golden/make_golden.py:c1_inputs:
    x uniform on [-1,1]^10 (seed 1), Genz oscillatory target (func/genz.py:32-54, shift 0.25, weights 1/(1+i)^2 on
    (x+1)/2) + 1e-2 N(0,1)."""
    rng = np.random.default_rng(1)
    x = rng.uniform(-1, 1, (n, d))
    w = 1.0 / (1 + np.arange(d)) ** 2
    noise = 1e-2 * rng.standard_normal(n)
    y = np.cos(2 * np.pi * 0.25 + ((x + 1) / 2) @ w) + noise
    return x, y


def golden_generator():
    """tests/golden/make_golden.py as a module (loaded by path; importing it does not import the reference): the
    tests regenerate the synthetic inputs of the larger goldens with the very functions that made them."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("pcq_make_golden", os.path.join(GOLDEN, "make_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod

import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import nemo
    nemo.lib()  # builds oracle/_build/libnem_oracle.so on first use
    return nemo


@pytest.fixture(scope="session")
def synth():
    from pangenomenem_b200 import synth as s
    return s


def make_case(n, d, seed=42, graph="pangenome", weighted=True):
    from pangenomenem_b200 import synth as s
    return s.make_pangenome(n, d, seed=seed, graph=graph, weighted=weighted)


@pytest.fixture(scope="session")
def small_pg():
    return make_case(3000, 70, seed=7)


@pytest.fixture(scope="session")
def engine():
    from pangenomenem_b200 import capi
    e = capi.Engine()
    yield e
    e.close()


def rel_close(a, b, rtol, atol=0.0):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    both_inf = np.isinf(a) & np.isinf(b) & (np.sign(a) == np.sign(b))
    with np.errstate(invalid="ignore"):
        ok = both_inf | (np.abs(a - b) <= atol + rtol * np.abs(b))
    return bool(np.all(ok))

import importlib
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
def pkg():
    """The product package (directory name has a hyphen)."""
    return importlib.import_module("rapidcfd-dev_b200")


@pytest.fixture(scope="session")
def meshmod():
    return importlib.import_module("rapidcfd-dev_b200.mesh")


@pytest.fixture(scope="session")
def orc():
    from oracle import ldu_oracle
    ldu_oracle.build()
    return ldu_oracle


def dense_from_ldu(n, l, u, diag, upper, lower=None):
    A = np.zeros((n, n))
    A[np.arange(n), np.arange(n)] = diag
    lower = upper if lower is None else lower
    for f in range(len(l)):
        A[l[f], u[f]] += upper[f]
        A[u[f], l[f]] += lower[f]
    return A

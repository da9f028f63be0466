import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


PRIOR_NAMES = ["mean:tcen", "mean:b", "mean:vel", "mean:RpRs", "mean:LD1", "mean:LD2"]
_KINDS = {1: "gaussian", 2: "normlim", 3: "evans"}


def priors_from_table(tab):
    """[6,6] prior table (lo,hi,kind,mu,sigma,pad) -> the reference's ordered priors dict."""
    d = {}
    for nm, row in zip(PRIOR_NAMES, np.asarray(tab)):
        d[nm] = [row[0], row[1]] if int(row[2]) == 0 else [row[0], row[1], _KINDS[int(row[2])], row[3], row[4]]
    return d


@pytest.fixture(scope="session")
def golden():
    return load_golden

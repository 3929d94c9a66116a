import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def engine():
    """The product path: in-tree libosp_b200.so on cuda:0.  No fallback: missing library or GPU is an error."""
    import outerspace_b200 as osp

    eng = osp.Engine(0)
    yield eng
    eng.close()


def load_npz(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))

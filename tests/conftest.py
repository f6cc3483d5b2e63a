import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden_cases():
    import glob
    d = os.path.join(REPO, "tests", "golden")
    # a case is an .npz with its model description next to it (c3_am_truth.npz is a table of exact values)
    return sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(d, "*.npz"))
                  if os.path.exists(f[:-4] + ".json"))


def load_golden(name):
    import numpy as np
    from astroemperor_b200.modelspec import ModelSpec
    d = os.path.join(REPO, "tests", "golden")
    g = np.load(os.path.join(d, name + ".npz"))
    spec = ModelSpec.from_json(open(os.path.join(d, name + ".json")).read())
    return g, spec


@pytest.fixture(scope="session")
def built_lib():
    """Path of the in-tree C-ABI library (built on demand; nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as g
    from astroemperor_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        g.build()
    return _lib.LIB_PATH

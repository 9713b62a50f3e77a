import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden_names(prefixes=("cc3", "win_", "md_", "ex_", "avg_")):
    out = []
    for f in sorted(glob.glob(os.path.join(GOLDEN, "*.npz"))):
        n = os.path.basename(f)[:-4]
        if n.startswith(tuple(prefixes)):
            out.append(n)
    return out


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


@pytest.fixture(scope="session")
def oracle():
    import pw_oracle
    return pw_oracle

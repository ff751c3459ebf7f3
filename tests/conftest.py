import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "rad-nerf_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


# the reference's op-by-op inference loop is test infrastructure (oracle/ops_frame.py): `model.render(..., path="ops")` in eval mode
# needs it registered
from oracle import ops_frame  # noqa: E402,F401


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """CPU oracle (test infrastructure).  Built on demand with gcc."""
    from oracle import oracle as o
    o.build()
    return o


def golden(name):
    import numpy as np
    import golden_cases as gc
    path = os.path.join(gc.GOLDEN_DIR, name + ".npz")
    if not os.path.exists(path):
        pytest.fail(f"golden fixture {path} missing: run oracle/make_golden.py on a GPU box and commit tests/golden/")
    return dict(np.load(path))

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def built():
    """Build the CUDA library (cross-compiles without a GPU) and the CPU oracle once per session."""
    import __graft_entry__ as g
    g.build_library()
    from oracle import binding as ob
    ob.build()
    return True


def golden_json(name):
    import json
    with open(os.path.join(GOLDEN, name)) as fh:
        return json.load(fh)


def golden_npz(name):
    import numpy as np
    return np.load(os.path.join(GOLDEN, name))

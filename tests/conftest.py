import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (oracle/oracle.py), built on demand.  Test infrastructure only."""
    import oracle as O
    O.build()
    O.set_threads(1)
    O.set_exp_mode(O.EXP_DET)
    return O


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN

import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
for p in (HERE, REPO):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle_built():
    import support
    if not os.path.exists(support.MINILMP_SO):
        support.build_oracle()
    return True


@pytest.fixture(scope="session")
def ctx():
    """One CUDA context for the session (GPU tests only)."""
    import lammps_plugins_b200 as b2
    c = b2.Context(0)
    yield c
    c.close()

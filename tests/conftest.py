import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


@pytest.fixture(scope="session")
def ctx():
    """One CUDA context for the whole GPU session.  No skip: a GPU test on a box without a
    usable device or without the built library must fail loudly."""
    import dusk_plonk_b200 as z
    c = z.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def cport():
    from oracle import cport as cp
    cp.build()
    return cp

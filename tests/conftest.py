import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionstart(session):
    """Build libgslift.so (nvcc, sm_100a) and the C oracle when they are missing or OLDER than
    any of their sources (both are git-ignored build products; the GPU box receives the prebuilt
    files together with the sources, so an up-to-date library is not rebuilt there)."""
    import importlib
    builder = importlib.import_module("3d_gaussian_splatting_project_b200.build")
    builder.build()                 # no-op unless stale()
    from oracle import oracle as orc
    orc.build()


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as orc
    orc.build()
    return orc

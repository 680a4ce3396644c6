import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_sessionstart(session):
    """Build what is missing or stale (nvcc cross-compiles without a GPU): the CUDA library and the C oracle."""
    from transrot_b200.build import build as build_cuda
    from oracle.oracle import build as build_oracle

    build_cuda()
    build_oracle()


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def dbase(golden_dir):
    from transrot_b200 import read_dbase

    return read_dbase(os.path.join(golden_dir, "shipped_species.dbase.txt"))


@pytest.fixture(scope="session")
def sample_cfg(golden_dir):
    from transrot_b200 import Config

    return Config.parseConfig(os.path.join(golden_dir, "sample8.config.txt"))


@pytest.fixture(scope="session")
def engine():
    """One context on cuda:0 for the whole GPU session; fails loudly when the library or the GPU is missing."""
    from transrot_b200 import Engine

    e = Engine(0)
    yield e
    e.close()

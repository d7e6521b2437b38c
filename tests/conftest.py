import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu)")


@pytest.fixture(scope="session")
def lsdb():
    from __graft_entry__ import load_package
    return load_package()


@pytest.fixture(scope="session")
def ctx(lsdb):
    c = lsdb.Context(0)
    yield c
    c.close()

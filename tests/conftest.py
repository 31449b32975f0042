import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def pkg():
    import __graft_entry__ as g
    return g.load_package()


@pytest.fixture(scope="session")
def corpus():
    import __graft_entry__ as g
    return g.load_submodule("corpus")


@pytest.fixture(scope="session")
def ora():
    import oracle_lib
    oracle_lib.oracle()
    return oracle_lib

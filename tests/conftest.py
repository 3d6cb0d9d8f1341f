import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def built():
    """The in-tree libraries (product .so and the oracle .so); built on demand."""
    from rays_b200 import _abi
    if not os.path.exists(_abi.lib_path()):
        import __graft_entry__ as ge
        ge.build()
    import _oracle
    _oracle.load()
    return True

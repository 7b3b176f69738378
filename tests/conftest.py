import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as o
    o.build()
    return o


@pytest.fixture(scope="session")
def mioc():
    import mioc_b200
    return mioc_b200


@pytest.fixture(scope="session")
def gpu_lib(mioc):
    """The CUDA library on a box with a GPU; GPU tests must never pass on a fallback."""
    import __graft_entry__ as g
    g.build()
    assert mioc.device_count() > 0, "no CUDA device visible: GPU tests cannot run"
    return mioc

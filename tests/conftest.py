import os
import sys
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def port():
    import orc
    orc.build("port")
    return orc.Oracle("port")


@pytest.fixture(scope="session")
def ref():
    """The compiled reference (oracle/_ref).  Built from /root/reference when that exists (dev container);
    on the GPU box only a prebuilt copy can be used."""
    import orc
    path = os.path.join(orc.ORACLE_DIR, "_ref", "libdabref.so")
    if os.path.isdir("/root/reference/src"):
        orc.build("ref")
    if not os.path.exists(path):
        pytest.skip("compiled reference (oracle/_ref) not available")
    return orc.Oracle("ref")

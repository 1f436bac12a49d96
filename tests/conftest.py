import os
import sys
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: takes more than a few seconds (full-size workloads)")


REF_SO = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libdabref.so")


def pytest_generate_tests(metafunc):
    """GPU parity tests run twice: against the C restatement ("port") and directly against the compiled
    reference classes ("ref", oracle/_ref/libdabref.so -- prebuilt in the dev container, travels to the GPU box)."""
    if "port" in metafunc.fixturenames and metafunc.definition.get_closest_marker("gpu"):
        kinds = ["port", "ref"] if (os.path.exists(REF_SO) or os.path.isdir("/root/reference/src")) else ["port"]
        metafunc.parametrize("port", kinds, indirect=True, scope="session")


@pytest.fixture(scope="session")
def port(request):
    import orc
    kind = getattr(request, "param", "port")
    if kind == "ref":
        if os.path.isdir("/root/reference/src"):
            orc.build("ref")
        if not os.path.exists(REF_SO):
            pytest.skip("compiled reference (oracle/_ref) not available")
        return orc.Oracle("ref")
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

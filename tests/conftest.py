import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def port():
    from _libs import Oracle
    return Oracle("port")


@pytest.fixture(scope="session")
def ref():
    from _libs import Oracle, ref_available
    if not ref_available():
        pytest.skip("oracle/_ref not built (needs /root/reference; prebuilt files travel with the snapshot)")
    return Oracle("ref")


@pytest.fixture(scope="session")
def avbd():
    import avbd_demo3d_b200 as m
    m.lib()
    return m

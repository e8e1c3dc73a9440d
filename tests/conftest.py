import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def sv():
    """The product package with its CUDA library loaded; fails loudly when the .so is missing."""
    import stitchingvideo_b200 as m
    m.lib()
    return m


@pytest.fixture(scope="session")
def gpu(sv):
    if sv.device_count() < 1:
        pytest.fail("no CUDA device visible: -m gpu tests must run on the GPU box (no CPU fallback exists)")
    return sv

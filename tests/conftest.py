import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    pyoracle.lib()
    return pyoracle


@pytest.fixture(scope="session")
def ggx_lut():
    """RGBA8 LUT rebuilt from the R,G fixture (tests/golden/make_ggx_lut_fixture.py)."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "ggx_lut_rg.npz"))
    rg = z["rg"]
    lut = np.zeros(rg.shape[:2] + (4,), dtype=np.uint8)
    lut[..., :2] = rg
    lut[..., 3] = 255
    return lut

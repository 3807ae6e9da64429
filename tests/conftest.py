import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def gs():
    """The product package with its CUDA library loaded; GPU tests fail loudly if the extension is missing."""
    import gaussiansplats3d_b200 as pkg
    pkg._native.load()
    if pkg._native.load().gs_device_count() <= 0:
        pytest.fail("no CUDA device visible to libgsplat_b200.so (gpu-marked test on a box without a GPU)")
    return pkg


os.environ.setdefault("OMP_NUM_THREADS", "8")

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built_artifacts():
    """Build the checker (oracle) and the product library once per session (nvcc cross-compiles
    without a GPU).  Building the oracle is not using it."""
    from oracle import raster as oracle_raster
    oracle_raster.build()
    from splatco_b200 import build as product_build
    product_build.build()
    yield

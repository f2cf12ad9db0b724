import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run on the GPU box with -m gpu")


@pytest.fixture(scope="session")
def oracle():
    from tests import oracle_lib
    return oracle_lib


@pytest.fixture(scope="session")
def product_lib():
    """the built CUDA library; building is part of the CPU-side check"""
    from maskfusion_b200 import build
    build.build()
    import maskfusion_b200
    return maskfusion_b200

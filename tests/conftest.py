import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

REFERENCE_DIR = "/root/reference"  # exists only in the build container, never on the GPU box


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: full-frame CPU renders (still part of the default CPU suite)")


@pytest.fixture(scope="session")
def oracle_lib():
    """Builds (if needed) and loads the CPU oracle — the checker, never the product."""
    from oracle import oracle

    oracle.build()
    return oracle.lib()


@pytest.fixture(scope="session")
def have_reference():
    return os.path.isdir(os.path.join(REFERENCE_DIR, "scenes"))

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def native_lib():
    """Build (if needed) and load the C-ABI library; CPU-safe."""
    from riemannian_motion_policies_b200 import build, _native
    build.build(verbose=False)
    return _native.lib()

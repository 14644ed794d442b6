import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def synthetic_urdfs():
    """The tests run on the mesh-free URDF fixtures: an explicit opt-in (URDFRobot has no silent fallback)."""
    import horopose_b200  # noqa: F401
    from horopose_b200 import synth
    synth.use_synthetic_urdfs()


@pytest.fixture(scope="session")
def hrp_lib():
    import horopose_b200  # noqa: F401
    from horopose_b200 import _lib
    return _lib.lib()

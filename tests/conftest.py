import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build librtcore_b200.so and the oracle once per session (no-ops when up to date)."""
    from raytracercore_b200 import build as B
    B.build()
    B.build_oracle()


SCENES = os.path.join(ROOT, "tests", "scenes")

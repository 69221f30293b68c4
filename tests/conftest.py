import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """libspx.so is git-ignored: on a fresh checkout the test session compiles it once (nvcc cross-compiles without a GPU).
    Building is not a fallback -- outside the tests a missing library is an error (_lib.lib())."""
    from self_play_reinforcement_learning_b200 import _lib, build
    if not os.environ.get("SPX_LIB_PATH") and not os.path.exists(_lib.LIB_PATH):
        build.build()

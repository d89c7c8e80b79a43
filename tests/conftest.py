import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden_files(variant=None):
    import glob
    pat = f"{variant}_B*.npz" if variant else "*_B*.npz"
    return sorted(glob.glob(os.path.join(GOLDEN, pat)))


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN

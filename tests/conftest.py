import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session", autouse=True)
def _oracle_library_is_current():
    """The host tests that link oracle/liboracle.so directly (the emulated-CTA tests) need it built from the current sources:
    the same `make` oracle/oracle_py.py runs on import (a no-op when it is up to date)."""
    import subprocess
    subprocess.call(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    yield

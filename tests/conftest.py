import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ORACLE_LIB = os.path.join(ROOT, "oracle", "libhived_oracle.so")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _build_oracle():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])


@pytest.fixture(scope="session")
def oracle_lib():
    """The CPU oracle behind the same C ABI (test infrastructure only)."""
    from hivedscheduler_b200 import _cabi
    _build_oracle()
    return _cabi.load_library(ORACLE_LIB)


@pytest.fixture(scope="session")
def cuda_lib():
    """The product library; GPU tests fail loudly if it is missing or no device is present."""
    from hivedscheduler_b200 import _cabi
    return _cabi.load_cuda_library()

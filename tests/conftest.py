import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def built_lib():
    """The in-tree libepc_b200.so (built on demand; nvcc cross-compiles without a GPU)."""
    import importlib
    import subprocess
    lib_mod = importlib.import_module("epc-net_b200._lib")
    if not os.path.exists(lib_mod.LIB_PATH):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "epc-net_b200", "csrc"), "-j8", "-s"])
    return lib_mod.load()

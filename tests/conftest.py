import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the oracle (gcc) and make sure libmeao.so exists (nvcc) before any test runs."""
    from oracle import oracle as oracle_mod
    oracle_mod.build()
    from miniengineao_b200 import build as build_mod
    if not os.path.exists(build_mod.LIB):
        build_mod.build()
    yield

import os
import sys

import pytest

# several band contexts share ONE GPU in the tests and handshake through spinning kernels: give every stream its own hardware
# queue so that one band's waiting exchange kernel can never sit in front of the neighbour's kernels (must be set before CUDA starts)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "run_last: collected after every other test (several spinning band contexts on ONE GPU: a "
                            "scheduling hazard of that arrangement must not stop a -x run before the other parity tests have reported)")


def _have_gpu() -> bool:
    try:
        import torch
        return bool(torch.cuda.is_available())
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """`gpu`-marked tests are SKIPPED (not failed) on a box without a GPU, so a plain `pytest tests/` is green on CPU."""
    items.sort(key=lambda it: 1 if "run_last" in it.keywords else 0)      # stable: everything else keeps its order
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="needs a B200 (no CUDA device visible); run with -m gpu on the GPU box")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the oracle (gcc) and libmeao.so (nvcc) before any test runs.  build() is a no-op when the library is newer
    than every source it depends on (build.is_stale), so a stale prebuilt .so can never be what the parity tests load."""
    from oracle import oracle as oracle_mod
    oracle_mod.build()
    from miniengineao_b200 import build as build_mod
    build_mod.build()
    yield

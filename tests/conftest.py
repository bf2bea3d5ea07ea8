"""pytest configuration: `-m gpu` tests need a B200 and call through the C ABI; everything else runs on CPU."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA sm_100a device (run with -m gpu on a B200)")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(autouse=True)
def _seeded():
    """Every test starts from the same RNG state (CPU and CUDA): tolerances are never exercised on a lucky/unlucky draw."""
    import torch

    torch.manual_seed(20260117)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(20260117)
    yield

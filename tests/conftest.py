"""pytest configuration: the ``gpu`` marker and sys.path for the repo root."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


def _has_cuda() -> bool:
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_cuda():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(autouse=True)
def _fresh_default_latent():
    """Flow's default latent is ONE instance shared by every Flow and latches its dim on first use (flow.py:20,
    distributions.py:31-32 - the reference behaves the same): un-latch it so that tests do not depend on their order."""
    try:
        from zenflow_b200 import flow as _flow
    except Exception:
        yield
        return
    _flow._DEFAULT_LATENT._Distribution__dim = None
    yield

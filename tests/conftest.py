"""pytest configuration: ``gpu`` marker, repo root on ``sys.path``.

``-m "not gpu"`` covers the oracle, host logic and the C-ABI load; ``-m gpu`` holds the parity
tests proper and runs on a B200.  ``/root/reference`` is never read by ``-m gpu`` tests.
"""

import os
import sys
from pathlib import Path

import pytest

# Two OpenMP runtimes in one process are a known crash source (reference tests/conftest.py:11-17).
os.environ.setdefault("OMP_NUM_THREADS", "1")

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 with -m gpu)")


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

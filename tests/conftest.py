"""pytest configuration: `gpu` marker + repo root on sys.path.

`-m "not gpu"` tests: oracle vs the committed golden vectors, host logic, and
that libpxr.so loads and exports every symbol include/pxr.h declares.
`-m gpu` tests: parity of the CUDA path (through the C ABI) against the oracle.
"""
import sys
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parent.parent
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)

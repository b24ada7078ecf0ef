import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
# The library reads HPTB_TUNE once; with it set, per-call development switches (HPTB_TUNE_NO_TMA, …) are honoured, which
# lets a test run the same call through two kernel paths (tests/test_tma_tile_gpu.py).  No switch is set by default.
os.environ.setdefault("HPTB_TUNE", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)

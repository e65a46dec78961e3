import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _has_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # a hung kernel blocks inside cudaDeviceSynchronize where no Python signal handler runs: the "thread" method
    # kills the process from a timer thread instead, so a deadlock fails the run instead of wedging the GPU box
    for item in items:
        if "gpu" in item.keywords:
            own = item.get_closest_marker("timeout")
            item.add_marker(pytest.mark.timeout(own.args[0] if own and own.args else 900, method="thread"))
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)

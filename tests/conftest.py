import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "two_gpus: needs two CUDA devices (NCCL ranks in separate processes + CUDA IPC); "
                                       "deselected on smaller boxes, where tests/test_gpu_sharded_one_device.py runs the "
                                       "same sharded path with all ranks on one device")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


HAVE_GPU = _have_gpu()


def _ngpu():
    if not HAVE_GPU:
        return 0
    import torch
    return torch.cuda.device_count()


def pytest_collection_modifyitems(config, items):
    if HAVE_GPU:
        if _ngpu() < 2:  # deselect (not skip): the one-device sharding tests cover the same code on this box
            drop = [it for it in items if "two_gpus" in it.keywords]
            if drop:
                config.hook.pytest_deselected(items=drop)
                items[:] = [it for it in items if "two_gpus" not in it.keywords]
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)

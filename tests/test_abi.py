"""CPU-side checks of the drop-in boundary: the shared library loads, exports exactly the
symbols include/megba_b200.h declares, the host-only partitioner works, and -- without a
GPU -- ba_create refuses loudly instead of falling back to a CPU path."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import HAVE_GPU, ROOT
from megba_b200 import capi


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "megba_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ba_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = capi.load()
    declared = _header_symbols()
    assert declared, "no declarations parsed from the header"
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/megba_b200.h but not exported"
    assert sorted(capi.SYMBOLS) == declared
    assert L.ba_abi_version() == 2


def test_partition_points_balances_edges_and_keeps_points_whole():
    rng = np.random.default_rng(0)
    deg = rng.integers(2, 40, size=5000)
    for ws in (1, 2, 3, 4, 8):
        cuts = capi.partition_points(deg, ws)
        assert cuts[0] == 0 and cuts[-1] == len(deg) and np.all(np.diff(cuts) >= 0)
        loads = np.array([deg[cuts[i]:cuts[i + 1]].sum() for i in range(ws)])
        assert loads.sum() == deg.sum()
        assert loads.max() - loads.min() <= 2 * deg.max()


def test_partition_points_degenerate():
    assert list(capi.partition_points(np.array([5]), 4)) in ([0, 0, 0, 0, 1], [0, 0, 0, 1, 1], [0, 0, 1, 1, 1], [0, 1, 1, 1, 1])
    assert list(capi.partition_points(np.zeros(0, np.int64), 2)) == [0, 0, 0]


@pytest.mark.skipif(HAVE_GPU, reason="only meaningful on a box without a GPU")
def test_create_without_gpu_fails_loudly():
    cam_idx = np.array([0, 1], np.int32)
    pt_idx = np.array([0, 0], np.int32)
    with pytest.raises(RuntimeError, match="no CPU fallback|-2"):
        capi.Solver(2, 1, cam_idx, pt_idx, np.zeros((2, 2)))


def test_create_rejects_bad_arguments():
    L = capi.load()
    h = C.c_void_p()
    opt = capi.Options(7, 0, 0, 0, 1)  # bad dtype
    assert L.ba_create(C.byref(opt), 1, 1, 0, None, None, None, C.byref(h)) == -1
    opt = capi.Options(0, 0, 0, 3, 2)  # rank >= world_size
    assert L.ba_create(C.byref(opt), 1, 1, 0, None, None, None, C.byref(h)) == -1
    assert L.ba_create(None, 1, 1, 0, None, None, None, C.byref(h)) == -1

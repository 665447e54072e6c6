"""The product's residual/Jacobian closed form (ba_geometry.cuh is __host__ __device__) instantiated on the host by a
test-only harness and compared with the oracle -- the no-GPU tier's check of the device arithmetic itself."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle.oracle import Oracle
from util import TOL, rel_rowmax, small_problem

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    out = tmp_path_factory.mktemp("geo") / "libgeo_host.so"
    src = os.path.join(HERE, "host_harness", "geometry_host.cu")
    env = dict(os.environ)
    env.pop("CC", None)
    subprocess.run(["/usr/local/cuda/bin/nvcc", "-O2", "-std=c++17", "-shared", "-Xcompiler", "-fPIC", "-ccbin", "/usr/bin/g++",
                    "-gencode", "arch=compute_100a,code=sm_100a", src, "-o", str(out)], check=True, env=env)
    return C.CDLL(str(out))


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_host_instantiation_of_device_geometry_matches_oracle(harness, dtype):
    P = small_problem()
    cam = P.cam.copy()
    cam[3, :3] = 1e-12  # one camera on the small-angle branch (analytical_derivatives.cu:131-158)
    O = Oracle(P.nC, P.nP, P.cam_idx, P.pt_idx, P.obs, dtype)
    r0, J0 = O.forward(cam, P.pt)
    ce = np.ascontiguousarray(cam[P.cam_idx], dtype)
    pe = np.ascontiguousarray(P.pt[P.pt_idx], dtype)
    ob = np.ascontiguousarray(P.obs, dtype)
    r = np.zeros((P.nE, 4), dtype)
    J = np.zeros((P.nE, 24), dtype)
    fn = harness.geo_rj_f64 if dtype == np.float64 else harness.geo_rj_f32
    fn(C.c_int64(P.nE), *(C.c_void_p(a.ctypes.data) for a in (ce, pe, ob, r, J)))
    tol = TOL[np.dtype(dtype)]["rj"]
    r0 = np.asarray(r0, np.float64).reshape(-1, 2)
    # same criteria as tests/test_gpu_parity.py::test_residual_and_jacobian (r is a difference of ~1e3-pixel numbers)
    bound = 50 * tol * max(1.0, np.abs(r0).max())
    assert np.abs(r[:, :2].astype(np.float64) - r0).max() <= bound
    assert np.abs(r[:, 2:].astype(np.float64) - r0).max() <= bound  # residual-only path (trial evaluation)
    assert rel_rowmax(J.reshape(-1, 2, 12), np.asarray(J0).reshape(-1, 2, 12)) <= 50 * tol

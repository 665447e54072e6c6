"""Fixed vertices (BaseVertex::fixed -- declared by the reference, include/vertex/base_vertex.h:34, never honoured by it):
the step of a solver with fixed cameras / points equals the solution of the damped normal equations restricted to the
free vertices (dense numpy reference built from the oracle's Jacobians), fixed parameters never move, and the C++ facade
forwards the flag."""
import numpy as np
import pytest

from megba_b200 import capi
from megba_b200.synthetic import make_synthetic
from oracle.oracle import Oracle
from util import rel_fro

pytestmark = pytest.mark.gpu


def _dense_reference(P, cam_fixed, pt_fixed, region):
    O = Oracle(P.nC, P.nP, P.cam_idx, P.pt_idx, P.obs, np.float64)
    r, J = O.forward(P.cam, P.pt)
    n = 9 * P.nC + 3 * P.nP
    Jd = np.zeros((2 * P.nE, n))
    for e in range(P.nE):
        c, p = P.cam_idx[e], P.pt_idx[e]
        Jd[2 * e:2 * e + 2, 9 * c:9 * c + 9] = J[e, :, :9]
        Jd[2 * e:2 * e + 2, 9 * P.nC + 3 * p:9 * P.nC + 3 * p + 3] = J[e, :, 9:]
    fixed = np.concatenate([np.repeat(cam_fixed.astype(bool), 9), np.repeat(pt_fixed.astype(bool), 3)])
    free = ~fixed
    Jf = Jd[:, free]
    H = Jf.T @ Jf
    g = -Jf.T @ r.ravel()
    H[np.diag_indices_from(H)] *= 1.0 + 1.0 / region  # processDiag, schur_LM_linear_system.cu:123-124
    dx = np.zeros(n)
    dx[free] = np.linalg.solve(H, g)
    return dx, fixed, O.cost(r)


@pytest.mark.parametrize("which", ["cameras", "points", "both"])
def test_fixed_vertices_solve_the_reduced_system(which):
    P = make_synthetic(8, 150, 700, seed=31)
    cam_fixed = np.zeros(P.nC, np.uint8)
    pt_fixed = np.zeros(P.nP, np.uint8)
    if which in ("cameras", "both"):
        cam_fixed[[0, 5]] = 1
    if which in ("points", "both"):
        pt_fixed[::7] = 1
    region = 1e3
    dx_ref, fixed, cost_ref = _dense_reference(P, cam_fixed, pt_fixed, region)
    S = capi.Solver(P.nC, P.nP, P.cam_idx, P.pt_idx, P.obs, dtype=np.float64)
    S.set_fixed(cam_fixed, pt_fixed)
    S.set_params(P.cam, P.pt)
    cost = S.linearize()
    assert abs(cost - cost_ref) <= 1e-10 * cost_ref
    S.process_diag(region)
    S.solve(max_iter=600, tol=1e-28, refuse_ratio=1e300)
    dx = S.read(capi.BUF_DX)
    assert np.all(dx[fixed] == 0.0)
    assert rel_fro(dx, dx_ref) <= 1e-8
    # the whole LM loop leaves fixed parameters bit-identical and still reduces the cost
    S.set_params(P.cam, P.pt)
    tr = S.lm_solve(lm_max_iter=5, initial_region=1e3, eps1=1e-6, eps2=1e-12, pcg_max_iter=100, pcg_tol=1e-10)
    cam, pt = S.get_params()
    assert np.array_equal(cam[cam_fixed.astype(bool)], P.cam[cam_fixed.astype(bool)])
    assert np.array_equal(pt[pt_fixed.astype(bool)], P.pt[pt_fixed.astype(bool)])
    assert tr[-1]["cost"] < tr[0]["cost"]
    moved = np.abs(cam[~cam_fixed.astype(bool)] - P.cam[~cam_fixed.astype(bool)]).max()
    assert moved > 0


def test_fixed_vertices_with_a_small_angle_camera_use_the_jacobian_stream():
    """The Jacobian-stream fallback (a camera with theta^2 <= eps) honours the flags too."""
    P = make_synthetic(8, 150, 700, seed=32)
    P.cam[3, :3] = 0.0
    cam_fixed = np.zeros(P.nC, np.uint8)
    cam_fixed[1] = 1
    pt_fixed = np.zeros(P.nP, np.uint8)
    pt_fixed[5::11] = 1
    region = 1e3
    dx_ref, fixed, _ = _dense_reference(P, cam_fixed, pt_fixed, region)
    S = capi.Solver(P.nC, P.nP, P.cam_idx, P.pt_idx, P.obs, dtype=np.float64)
    S.set_fixed(cam_fixed, pt_fixed)
    S.set_params(P.cam, P.pt)
    S.linearize()
    S.process_diag(region)
    S.solve(max_iter=600, tol=1e-28, refuse_ratio=1e300)
    dx = S.read(capi.BUF_DX)
    assert np.all(dx[fixed] == 0.0)
    assert rel_fro(dx, dx_ref) <= 1e-8

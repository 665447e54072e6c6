"""CPU tests of the oracle itself (it is unpinned upstream, so it is pinned here by
independent mathematics: finite differences, a dense solve, an independent projection)."""
import json
import os

import numpy as np
import pytest

from megba_b200.synthetic import make_synthetic, project, read_bal, write_bal
from oracle.oracle import Oracle
from util import medium_problem, rel_fro, small_problem

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _oracle(P, dtype=np.float64):
    return Oracle(P.nC, P.nP, P.cam_idx, P.pt_idx, P.obs, dtype)


def test_residual_matches_independent_projection():
    P = medium_problem()
    r, _ = _oracle(P).forward(P.cam, P.pt)
    ref = project(P.cam[P.cam_idx], P.pt[P.pt_idx]) - P.obs
    assert np.abs(r - ref).max() < 1e-9 * max(1.0, np.abs(ref).max())


def test_jacobian_central_differences():
    P = small_problem()
    O = _oracle(P)
    _, J = O.forward(P.cam, P.pt)

    def res(cam, pt):
        return O.forward(cam, pt, jac=False)[0]

    h = 1e-6
    Jfd = np.zeros_like(J)
    for j in range(9):
        cp, cm = P.cam.copy(), P.cam.copy()
        cp[:, j] += h
        cm[:, j] -= h
        Jfd[:, :, j] = (res(cp, P.pt) - res(cm, P.pt)) / (2 * h)
    for j in range(3):
        pp, pm = P.pt.copy(), P.pt.copy()
        pp[:, j] += h
        pm[:, j] -= h
        Jfd[:, :, 9 + j] = (res(P.cam, pp) - res(P.cam, pm)) / (2 * h)
    scale = np.abs(J).max(axis=(0, 1))
    assert (np.abs(Jfd - J).max(axis=(0, 1)) / scale).max() < 5e-7


def test_small_angle_branch_is_continuous():
    """theta^2 <= eps branch (analytical_derivatives.cu:131-158) agrees with the general branch nearby."""
    P = small_problem()
    cam_a, cam_b = P.cam.copy(), P.cam.copy()
    cam_a[:, :3] = 1e-9 * np.array([1.0, -2.0, 0.5])   # first-order branch
    cam_b[:, :3] = 1e-7 * np.array([1.0, -2.0, 0.5])   # general branch
    O = _oracle(P)
    ra, Ja = O.forward(cam_a, P.pt)
    rb, Jb = O.forward(cam_b, P.pt)
    assert np.abs(ra - rb).max() < 1e-3
    assert rel_fro(Ja, Jb) < 1e-5


def test_build_matches_dense_JtJ():
    P = small_problem()
    O = _oracle(P)
    r, J = O.forward(P.cam, P.pt)
    Hpp, Hll, W, g = O.build(r, J)
    nX = 9 * P.nC + 3 * P.nP
    Jd = np.zeros((2 * P.nE, nX))
    for e in range(P.nE):
        c, p = P.cam_idx[e], P.pt_idx[e]
        Jd[2 * e:2 * e + 2, 9 * c:9 * c + 9] = J[e, :, :9]
        Jd[2 * e:2 * e + 2, 9 * P.nC + 3 * p:9 * P.nC + 3 * p + 3] = J[e, :, 9:]
    H = Jd.T @ Jd
    for c in range(P.nC):
        assert np.allclose(Hpp[c], H[9 * c:9 * c + 9, 9 * c:9 * c + 9], rtol=1e-10, atol=1e-6)
    for p in range(0, P.nP, 37):
        o = 9 * P.nC + 3 * p
        assert np.allclose(Hll[p], H[o:o + 3, o:o + 3], rtol=1e-10, atol=1e-8)
    e = 123
    c, p = P.cam_idx[e], P.pt_idx[e]
    assert np.allclose(W[e], H[9 * c:9 * c + 9, 9 * P.nC + 3 * p:9 * P.nC + 3 * p + 3], rtol=1e-10, atol=1e-8)
    assert np.allclose(g, -Jd.T @ r.ravel(), rtol=1e-10, atol=1e-6)


@pytest.mark.parametrize("implicit", [False, True])
def test_pcg_converges_to_dense_solution(implicit):
    P = small_problem()
    O = _oracle(P)
    r, J = O.forward(P.cam, P.pt)
    Hpp, Hll, W, g = O.build(r, J)
    region = 1e3
    HppD, _ = O.process_diag(Hpp, region)
    HllD, _ = O.process_diag(Hll, region)
    dx, it, done, _ = O.schur_pcg(HppD, HllD, None if implicit else W, J if implicit else None, g,
                                  np.zeros(9 * P.nC + 3 * P.nP), 500, 1e-26, 1e30)
    nX = 9 * P.nC + 3 * P.nP
    H = np.zeros((nX, nX))
    for c in range(P.nC):
        H[9 * c:9 * c + 9, 9 * c:9 * c + 9] = HppD[c]
    for p in range(P.nP):
        o = 9 * P.nC + 3 * p
        H[o:o + 3, o:o + 3] = HllD[p]
    for e in range(P.nE):
        c, p = P.cam_idx[e], P.pt_idx[e]
        o = 9 * P.nC + 3 * p
        H[9 * c:9 * c + 9, o:o + 3] += W[e]
        H[o:o + 3, 9 * c:9 * c + 9] += W[e].T
    ref = np.linalg.solve(H, g)
    assert rel_fro(dx, ref) < 1e-7


def test_block_inverse():
    rng = np.random.default_rng(0)
    O = _oracle(small_problem())
    for n in (3, 9):
        M = rng.normal(size=(50, 2 * n, n))
        A = np.einsum("bki,bkj->bij", M, M)
        inv = O.block_inverse(A)
        assert np.abs(np.einsum("bij,bjk->bik", A, inv) - np.eye(n)).max() < 1e-9


def test_lm_cost_monotone_and_f32_tracks_f64():
    P = medium_problem()
    x, k, tr = _oracle(P).lm_solve(P.x(), lm_max_iter=8, initial_region=1e4)
    costs = [t["cost"] for t in tr]
    assert all(b <= a for a, b in zip(costs, costs[1:]))
    assert costs[-1] < 0.05 * costs[0]
    x32, k32, tr32 = _oracle(P, np.float32).lm_solve(P.x(), lm_max_iter=8, initial_region=1e4)
    assert abs(tr32[1]["cost"] - tr[1]["cost"]) < 2e-3 * tr[1]["cost"]


def test_golden_oracle_trace():
    """Regression pin: tests/golden/oracle_lm_small.json was produced by tests/golden/make_golden.py from
    this oracle (NOT from upstream -- upstream has no fixtures).  Guards the oracle against drift."""
    with open(os.path.join(GOLD, "oracle_lm_small.json")) as fh:
        gold = json.load(fh)
    P = make_synthetic(*gold["shape"], seed=gold["seed"])
    _, _, tr = _oracle(P).lm_solve(P.x(), **gold["params"])
    assert [t["accepted"] for t in tr] == gold["accepted"]
    assert [t["pcg_iters"] for t in tr] == gold["pcg_iters"]
    assert np.allclose([t["cost"] for t in tr], gold["cost"], rtol=1e-9)


def test_bal_text_roundtrip(tmp_path):
    P = make_synthetic(5, 40, 130, seed=9)
    path = str(tmp_path / "problem-5-40-pre.txt")
    write_bal(path, P)
    Q = read_bal(path)
    assert (Q.nC, Q.nP, Q.nE) == (P.nC, P.nP, P.nE)
    assert np.array_equal(Q.cam_idx, P.cam_idx) and np.array_equal(Q.pt_idx, P.pt_idx)
    assert np.array_equal(Q.obs, P.obs) and np.array_equal(Q.cam, P.cam) and np.array_equal(Q.pt, P.pt)

"""GPU parity tests: the CUDA path, called through the C-ABI (megba_b200.capi -> libmegba_b200.so),
against the CPU oracle on the same seeded inputs.  Tolerances: util.TOL."""
import numpy as np
import pytest

from megba_b200 import capi
from megba_b200.synthetic import make_named, make_synthetic
from oracle.oracle import Oracle
from util import TOL, bigpoint_problem, medium_problem, rel_fro, rel_rowmax, small_problem

pytestmark = pytest.mark.gpu

DTYPES = [np.float64, np.float32]
PROBLEMS = {"small": small_problem, "medium": medium_problem, "bigpoint": bigpoint_problem}


def _pair(P, dtype, implicit=False):
    S = capi.Solver(P.nC, P.nP, P.cam_idx, P.pt_idx, P.obs, dtype=dtype, implicit=implicit)
    S.set_params(P.cam, P.pt)
    O = Oracle(P.nC, P.nP, P.cam_idx, P.pt_idx, P.obs, dtype)
    return S, O


def _oracle_system(O, P, region, implicit=False, info=None):
    r, J = O.forward(P.cam, P.pt)
    Hpp, Hll, W, g = O.build(r, J, info=info, explicit=not implicit)
    HppD, _ = O.process_diag(Hpp, region)
    HllD, _ = O.process_diag(Hll, region)
    return r, J, Hpp, Hll, W, g, HppD, HllD


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("name", ["small", "medium", "bigpoint"])
def test_residual_and_jacobian(name, dtype):
    P = PROBLEMS[name]()
    if name == "bigpoint":
        P.cam[3, :3] = 0.0           # exact zero rotation -> first-order branch
        P.cam[4, :3] = [1e-9, -2e-9, 5e-10]
    S, O = _pair(P, dtype)
    r, J = S.eval_edges()
    r0, J0 = O.forward(P.cam, P.pt)
    t = TOL[np.dtype(dtype)]["rj"]
    assert np.abs(r.astype(np.float64) - r0).max() <= t * max(1.0, np.abs(r0).max()) * 50
    assert rel_rowmax(J, J0) <= t * 50


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("name", ["small", "medium", "bigpoint"])
def test_linearize_explicit(name, dtype):
    P = PROBLEMS[name]()
    S, O = _pair(P, dtype)
    cost = S.linearize()
    r, J = O.forward(P.cam, P.pt)
    Hpp, Hll, W, g = O.build(r, J)
    t = TOL[np.dtype(dtype)]
    assert abs(cost - O.cost(r)) <= t["cost"] * O.cost(r)
    assert rel_fro(S.read(capi.BUF_HPP), Hpp) <= t["H"]
    assert rel_fro(S.read(capi.BUF_HLL), Hll) <= t["H"]
    assert rel_fro(S.read(capi.BUF_W), W) <= t["H"]
    assert rel_fro(S.read(capi.BUF_G), g) <= t["H"]
    # per-block check so that a single wrong small block cannot hide in the Frobenius norm
    assert rel_rowmax(S.read(capi.BUF_HLL).reshape(P.nP, 9), Hll.reshape(P.nP, 9)) <= 1e3 * t["H"]
    assert rel_rowmax(S.read(capi.BUF_W).reshape(P.nE, 27), W.reshape(P.nE, 27)) <= 1e3 * t["H"]


@pytest.mark.parametrize("dtype", DTYPES)
def test_linearize_implicit(dtype):
    P = medium_problem()
    S, O = _pair(P, dtype, implicit=True)
    cost = S.linearize()
    r, J = O.forward(P.cam, P.pt)
    Hpp, Hll, _, g = O.build(r, J, explicit=False)
    t = TOL[np.dtype(dtype)]
    assert abs(cost - O.cost(r)) <= t["cost"] * O.cost(r)
    assert rel_fro(S.read(capi.BUF_HPP), Hpp) <= t["H"]
    assert rel_fro(S.read(capi.BUF_HLL), Hll) <= t["H"]
    assert rel_fro(S.read(capi.BUF_G), g) <= t["H"]
    assert rel_rowmax(S.read(capi.BUF_J).reshape(P.nE * 2, 12), J.reshape(P.nE * 2, 12)) <= 50 * t["rj"]


@pytest.mark.parametrize("dtype", DTYPES)
def test_linearize_with_information(dtype):
    P = small_problem()
    rng = np.random.default_rng(11)
    a = rng.uniform(0.5, 2.0, size=P.nE)
    b = rng.uniform(-0.3, 0.3, size=P.nE)
    c = rng.uniform(0.5, 2.0, size=P.nE)
    info = np.stack([a, b, b, c], axis=1)  # symmetric positive definite 2x2
    S, O = _pair(P, dtype)
    S.set_information(info)
    S.linearize()
    r, J = O.forward(P.cam, P.pt)
    Hpp, Hll, W, g = O.build(r, J, info=info)
    t = TOL[np.dtype(dtype)]
    assert rel_fro(S.read(capi.BUF_HPP), Hpp) <= t["H"]
    assert rel_fro(S.read(capi.BUF_HLL), Hll) <= t["H"]
    assert rel_fro(S.read(capi.BUF_W), W) <= t["H"]
    assert rel_fro(S.read(capi.BUF_G), g) <= t["H"]
    S.set_information(None)
    S.linearize()
    Hpp0, _, _, _ = O.build(r, J)
    assert rel_fro(S.read(capi.BUF_HPP), Hpp0) <= t["H"]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("implicit", [False, True])
@pytest.mark.parametrize("name", ["medium", "bigpoint"])
def test_schur_operator(name, implicit, dtype):
    P = PROBLEMS[name]()
    S, O = _pair(P, dtype, implicit=implicit)
    S.linearize()
    region = 1e3
    S.process_diag(region)
    r, J, Hpp, Hll, W, g, HppD, HllD = _oracle_system(O, P, region, implicit)
    p = np.random.default_rng(3).normal(size=9 * P.nC)
    q = S.schur_apply(p)
    q0 = O.schur_mul(HppD, O.block_inverse(HllD), W, J if implicit else None, p)
    assert rel_fro(q, q0) <= 10 * TOL[np.dtype(dtype)]["H"]
    # linearity and symmetry of S (size-independent properties, also used at full size)
    p2 = np.random.default_rng(4).normal(size=9 * P.nC)
    q2 = S.schur_apply(p2)
    q12 = S.schur_apply(0.5 * p - 2.0 * p2)
    eps = 1e-10 if dtype == np.float64 else 2e-4
    assert rel_fro(q12, 0.5 * q.astype(np.float64) - 2.0 * q2) <= eps
    assert abs(p2 @ q.astype(np.float64) - p @ q2.astype(np.float64)) <= eps * abs(p @ q.astype(np.float64))


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("implicit", [False, True])
@pytest.mark.parametrize("name", ["small", "medium", "bigpoint"])
def test_schur_pcg_fixed_iterations(name, implicit, dtype):
    """Step vector after a FIXED number of PCG iterations (tol = 0, refuse = inf) incl. back-substitution."""
    P = PROBLEMS[name]()
    S, O = _pair(P, dtype, implicit=implicit)
    S.linearize()
    region = 1e3
    S.process_diag(region)
    r, J, Hpp, Hll, W, g, HppD, HllD = _oracle_system(O, P, region, implicit)
    n_it = 12
    it, done = S.solve(max_iter=n_it, tol=0.0, refuse_ratio=1e300)
    dx0, it0, done0, _ = O.schur_pcg(HppD, HllD, W, J if implicit else None, g, np.zeros(9 * P.nC + 3 * P.nP),
                                     n_it, 0.0, 1e300)
    assert it == it0 == n_it and done == done0
    dx = S.read(capi.BUF_DX)
    t = TOL[np.dtype(dtype)]["dx"]
    assert rel_fro(dx[:9 * P.nC], dx0[:9 * P.nC]) <= t
    assert rel_fro(dx[9 * P.nC:], dx0[9 * P.nC:]) <= t


@pytest.mark.parametrize("dtype", DTYPES)
def test_schur_pcg_reference_stopping_rules_and_warm_start(dtype):
    """README parameters (tol 1e-1, refuse 1): same iteration count, same `done`, same step; then a warm start."""
    P = medium_problem()
    S, O = _pair(P, dtype)
    S.linearize()
    region = 1e4
    S.process_diag(region)
    r, J, Hpp, Hll, W, g, HppD, HllD = _oracle_system(O, P, region)
    it, done = S.solve(max_iter=100, tol=1e-1, refuse_ratio=1.0)
    dx0, it0, done0, _ = O.schur_pcg(HppD, HllD, W, None, g, np.zeros(9 * P.nC + 3 * P.nP), 100, 1e-1, 1.0)
    assert (it, done) == (it0, done0)
    t = TOL[np.dtype(dtype)]["dx"]
    assert rel_fro(S.read(capi.BUF_DX), dx0) <= t
    # warm start from a perturbed step (x0 != 0 path: r0 = v - S x0)
    x0 = dx0 * 0.9
    S.write_dx(x0)
    it, done = S.solve(max_iter=5, tol=0.0, refuse_ratio=1e300)
    dx1, it1, _, _ = O.schur_pcg(HppD, HllD, W, None, g, x0, 5, 0.0, 1e300)
    assert it == it1 == 5
    assert rel_fro(S.read(capi.BUF_DX), dx1) <= t


def test_pcg_refuse_rule_restores_previous_iterate():
    """refuse_ratio < 1 makes rho > refuse*rho_min fire early (schur_pcg_solver.cu:288-295)."""
    P = medium_problem()
    S, O = _pair(P, np.float64)
    S.linearize()
    S.process_diag(1e3)
    r, J, Hpp, Hll, W, g, HppD, HllD = _oracle_system(O, P, 1e3)
    it, done = S.solve(max_iter=50, tol=0.0, refuse_ratio=0.05)
    dx0, it0, done0, _ = O.schur_pcg(HppD, HllD, W, None, g, np.zeros(9 * P.nC + 3 * P.nP), 50, 0.0, 0.05)
    assert (it, done) == (it0, done0) and 0 < it < 50
    assert rel_fro(S.read(capi.BUF_DX), dx0) <= 1e-8


@pytest.mark.parametrize("dtype", DTYPES)
def test_lm_scalars(dtype):
    P = medium_problem()
    S, O = _pair(P, dtype)
    F = S.linearize()
    S.process_diag(1e3)
    S.solve(max_iter=10, tol=0.0, refuse_ratio=1e300)
    dx = S.read(capi.BUF_DX).astype(np.float64)
    dxn, xn = S.step_norms()
    t = TOL[np.dtype(dtype)]["cost"]
    assert abs(dxn - np.linalg.norm(dx)) <= 10 * t * np.linalg.norm(dx)
    assert abs(xn - np.linalg.norm(P.x())) <= 10 * t * np.linalg.norm(P.x())
    cost_new, model_sq = S.eval_trial()
    r, J = O.forward(P.cam, P.pt)
    dxT = dx.astype(dtype)
    xt = P.x(dtype) + dxT
    r2, _ = O.forward(xt[:9 * P.nC], xt[9 * P.nC:], jac=False)
    assert abs(cost_new - O.cost(r2)) <= 10 * t * O.cost(r2)
    m0 = O.jdx_plus_r_sq(J, r, dxT)
    assert abs(model_sq - m0) <= 10 * t * m0
    g = S.read(capi.BUF_G).astype(np.float64)
    assert abs(S.grad_inf_norm() - np.abs(g).max()) <= 1e-12 * np.abs(g).max()
    # accept: x += dx ; reject: dx <- backup
    S.accept_step()
    cam, pt = S.get_params()
    assert rel_fro(np.concatenate([cam.ravel(), pt.ravel()]), xt) <= (1e-15 if dtype == np.float64 else 1e-6)


@pytest.mark.parametrize("dtype,implicit", [(np.float64, False), (np.float64, True), (np.float32, False)])
def test_lm_trace_matches_oracle(dtype, implicit):
    """End-to-end LM (lm_algo.cu:139-223): per-iteration cost, accept/reject sequence, PCG iteration counts.
    tau is chosen so that the run contains rejected steps as well."""
    P = make_synthetic(30, 2000, 9000, seed=5)
    # perturb harder so the first steps are genuinely non-linear
    rng = np.random.default_rng(1)
    P.cam[:, :3] += rng.normal(0, 0.3, size=(P.nC, 3))
    P.cam[:, 3:6] += rng.normal(0, 2.0, size=(P.nC, 3))
    P.pt += rng.normal(0, 3.0, size=P.pt.shape)
    S, O = _pair(P, dtype, implicit=implicit)
    kw = dict(lm_max_iter=12, initial_region=1e8, eps1=1e-3, eps2=1e-12, pcg_max_iter=60, pcg_tol=1e-6,
              pcg_refuse=1.0)
    tr = S.lm_solve(**kw)
    x0, k0, tr0 = O.lm_solve(P.x(), implicit=implicit, **kw)
    t = TOL[np.dtype(dtype)]["cost"]
    n = min(len(tr), len(tr0))
    # compare while the iteration is still making real progress (beyond that accept/reject is rounding noise)
    c0 = np.array([r["cost"] for r in tr0])
    sig = [i for i in range(n) if i == 0 or abs(c0[i - 1] - c0[i]) > 1e-6 * c0[i] or tr0[i]["accepted"] == 0]
    last = max(i for i in sig) if dtype == np.float64 else 3
    assert last >= 3
    for i in range(last + 1):
        assert tr[i]["accepted"] == tr0[i]["accepted"], (i, tr, tr0)
        assert abs(tr[i]["cost"] - tr0[i]["cost"]) <= 10 * t * tr0[i]["cost"], (i, tr[i], tr0[i])
        if dtype == np.float64:
            assert tr[i]["pcg_iters"] == tr0[i]["pcg_iters"], (i, tr[i], tr0[i])
            # gain ratio and trust region (lm_algo.cu:181-198): inside ba_lm_solve the model decrease comes from the
            # quadratic form dx^T H dx - 2 g^T dx evaluated in the back-substitution pass, not from an edge pass
            if i > 0:
                assert abs(tr[i]["rho"] - tr0[i]["rho"]) <= 1e-6 * max(1.0, abs(tr0[i]["rho"])), (i, tr[i], tr0[i])
                assert abs(tr[i]["region"] - tr0[i]["region"]) <= 1e-5 * tr0[i]["region"], (i, tr[i], tr0[i])
    if dtype == np.float64:
        cam, pt = S.get_params()
        assert rel_fro(np.concatenate([cam.ravel(), pt.ravel()]), x0) <= 1e-6


def test_golden_lm_trace_fp64():
    """Same fixture as tests/test_oracle.py::test_golden_oracle_trace, now through the CUDA path."""
    import json
    import os
    with open(os.path.join(os.path.dirname(__file__), "golden", "oracle_lm_small.json")) as fh:
        gold = json.load(fh)
    P = make_synthetic(*gold["shape"], seed=gold["seed"])
    S, _ = _pair(P, np.float64)
    tr = S.lm_solve(**gold["params"])
    assert [t["accepted"] for t in tr][:3] == gold["accepted"][:3]
    assert [t["pcg_iters"] for t in tr][:3] == gold["pcg_iters"][:3]
    assert np.allclose([t["cost"] for t in tr][:3], gold["cost"][:3], rtol=1e-8)


def test_tiny_and_ragged_graphs():
    """Smallest legal graph, points with a single observation, unsorted input order."""
    rng = np.random.default_rng(0)
    P = make_synthetic(3, 6, 14, seed=4)
    perm = rng.permutation(P.nE)
    S = capi.Solver(P.nC, P.nP, P.cam_idx[perm], P.pt_idx[perm], P.obs[perm])
    S.set_params(P.cam, P.pt)
    O = Oracle(P.nC, P.nP, P.cam_idx[perm], P.pt_idx[perm], P.obs[perm])
    cost = S.linearize()
    r, J = O.forward(P.cam, P.pt)
    Hpp, Hll, W, g = O.build(r, J)
    assert abs(cost - O.cost(r)) <= 1e-10 * O.cost(r)
    assert rel_fro(S.read(capi.BUF_W), W) <= 1e-10 and rel_fro(S.read(capi.BUF_HPP), Hpp) <= 1e-10
    # drop observations so that some points keep exactly one
    keep = np.ones(P.nE, bool)
    for p in (1, 4):
        idx = np.flatnonzero(P.pt_idx == p)
        keep[idx[1:]] = False
    S2 = capi.Solver(P.nC, P.nP, P.cam_idx[keep], P.pt_idx[keep], P.obs[keep])
    S2.set_params(P.cam, P.pt)
    O2 = Oracle(P.nC, P.nP, P.cam_idx[keep], P.pt_idx[keep], P.obs[keep])
    S2.linearize()
    r2, J2 = O2.forward(P.cam, P.pt)
    Hpp2, Hll2, W2, g2 = O2.build(r2, J2)
    assert rel_fro(S2.read(capi.BUF_HLL), Hll2) <= 1e-10 and rel_fro(S2.read(capi.BUF_G), g2) <= 1e-10


def test_call_order_is_enforced():
    P = small_problem()
    S = capi.Solver(P.nC, P.nP, P.cam_idx, P.pt_idx, P.obs)
    S.set_params(P.cam, P.pt)
    with pytest.raises(RuntimeError, match="before ba_linearize"):
        S.solve()


@pytest.mark.timeout(900)
def test_full_size_venice_properties_and_parity():
    """BASELINE.json configs[2] size (Venice-1778 shape, synthetic): direct parity of the linearisation
    and of a fixed-count PCG solve against the oracle, plus size-independent operator properties."""
    P = make_named("venice-1778")
    S, O = _pair(P, np.float64)
    cost = S.linearize()
    r, J = O.forward(P.cam, P.pt)
    Hpp, Hll, W, g = O.build(r, J)
    assert abs(cost - O.cost(r)) <= 1e-9 * O.cost(r)
    assert rel_fro(S.read(capi.BUF_HPP), Hpp) <= 1e-10
    assert rel_fro(S.read(capi.BUF_G), g) <= 1e-10
    assert rel_fro(S.read(capi.BUF_HLL), Hll) <= 1e-10
    region = 1e4
    S.process_diag(region)
    rng = np.random.default_rng(0)
    p1, p2 = rng.normal(size=9 * P.nC), rng.normal(size=9 * P.nC)
    q1, q2 = S.schur_apply(p1), S.schur_apply(p2)
    assert abs(p2 @ q1 - p1 @ q2) <= 1e-9 * abs(p1 @ q1)          # S symmetric
    assert p1 @ q1 > 0 and p2 @ q2 > 0                             # S positive definite
    q12 = S.schur_apply(p1 + 3.0 * p2)
    assert rel_fro(q12, q1 + 3.0 * q2) <= 1e-10                    # linear
    HppD, _ = O.process_diag(Hpp, region)
    HllD, _ = O.process_diag(Hll, region)
    it, _ = S.solve(max_iter=8, tol=0.0, refuse_ratio=1e300)
    dx0, it0, _, _ = O.schur_pcg(HppD, HllD, W, None, g, np.zeros(9 * P.nC + 3 * P.nP), 8, 0.0, 1e300)
    assert it == it0 == 8
    assert rel_fro(S.read(capi.BUF_DX), dx0) <= 1e-8


@pytest.mark.parametrize("workload,dtype,implicit", [("ladybug-49", np.float64, False),
                                                     ("trafalgar-257", np.float64, False),
                                                     ("venice-1778", np.float64, True),
                                                     ("final-13682", np.float32, False)])
def test_full_size_baseline_configs(workload, dtype, implicit):
    """The other BASELINE.json configurations at their full sizes (synthetic problems of the BAL shapes): configs[0]
    Ladybug-49 fp64, configs[1] Trafalgar-257 fp64, configs[4] Venice-1778 matrix-free fp64 and configs[3]
    Final-13682 fp32 (the size that takes the cooperative-grid camera kernel): linearisation + 8 fixed PCG iterations
    against the oracle, plus operator symmetry / linearity."""
    P = make_named(workload)
    S, O = _pair(P, dtype, implicit=implicit)
    t = TOL[np.dtype(dtype)]
    cost = S.linearize()
    r, J = O.forward(P.cam, P.pt)
    Hpp, Hll, W, g = O.build(r, J, explicit=not implicit)
    assert abs(cost - O.cost(r)) <= t["cost"] * O.cost(r)
    assert rel_fro(S.read(capi.BUF_HPP), Hpp) <= t["H"]
    assert rel_fro(S.read(capi.BUF_HLL), Hll) <= t["H"]
    assert rel_fro(S.read(capi.BUF_G), g) <= t["H"]
    region = 1e4
    S.process_diag(region)
    rng = np.random.default_rng(0)
    p1, p2 = rng.normal(size=9 * P.nC), rng.normal(size=9 * P.nC)
    q1, q2 = S.schur_apply(p1).astype(np.float64), S.schur_apply(p2).astype(np.float64)
    eps = 1e-9 if dtype == np.float64 else 5e-4
    assert abs(p2 @ q1 - p1 @ q2) <= eps * abs(p1 @ q1)
    assert p1 @ q1 > 0 and p2 @ q2 > 0
    assert rel_fro(S.schur_apply(p1 + 3.0 * p2), q1 + 3.0 * q2) <= eps
    HppD, _ = O.process_diag(Hpp, region)
    HllD, _ = O.process_diag(Hll, region)
    del r
    it, _ = S.solve(max_iter=8, tol=0.0, refuse_ratio=1e300)
    dx0, it0, _, _ = O.schur_pcg(HppD, HllD, W, J if implicit else None, g, np.zeros(9 * P.nC + 3 * P.nP), 8, 0.0, 1e300)
    assert it == it0 == 8
    assert rel_fro(S.read(capi.BUF_DX), dx0) <= t["dx"]


@pytest.mark.parametrize("dtype", DTYPES)
def test_many_cameras_cooperative_grid_pcg_kernel(dtype):
    """3000 cameras do not fit one 16-CTA cluster at <= 2 rows per thread: the camera-space PCG kernel runs as a
    cooperative grid (global partials + grid barrier) instead.  Same checks as for the cluster flavour."""
    P = make_synthetic(3000, 40000, 300000, seed=17)  # >= 13 observations per camera: every Hpp block is well posed
    S, O = _pair(P, dtype)
    S.linearize()
    region = 1e3
    S.process_diag(region)
    r, J, Hpp, Hll, W, g, HppD, HllD = _oracle_system(O, P, region)
    it, done = S.solve(max_iter=10, tol=0.0, refuse_ratio=1e300)
    dx0, it0, done0, _ = O.schur_pcg(HppD, HllD, W, None, g, np.zeros(9 * P.nC + 3 * P.nP), 10, 0.0, 1e300)
    assert (it, done) == (it0, done0)
    assert rel_fro(S.read(capi.BUF_DX), dx0) <= TOL[np.dtype(dtype)]["dx"]
    it, done = S.solve(max_iter=100, tol=1e-1, refuse_ratio=1.0)   # warm start + reference stopping rules
    dx1, it1, done1, _ = O.schur_pcg(HppD, HllD, W, None, g, dx0, 100, 1e-1, 1.0)
    if dtype == np.float64:
        assert (it, done) == (it1, done1)
    assert rel_fro(S.read(capi.BUF_DX), dx1) <= 10 * TOL[np.dtype(dtype)]["dx"]


def test_empty_graph_and_unobserved_vertices():
    """No observations at all, and vertices without observations inside a normal graph: nothing crashes, the unobserved
    blocks stay zero and the observed part matches the oracle run on the same (ragged) input."""
    S = capi.Solver(2, 3, np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros((0, 2)))
    S.set_params(np.zeros((2, 9)), np.zeros((3, 3)))
    assert S.linearize() == 0.0
    it, done = S.solve(max_iter=5, tol=1e-1, refuse_ratio=1.0)
    assert np.all(S.read(capi.BUF_DX) == 0.0)
    # camera 5 and points 7, 8 have no observations
    P = make_synthetic(5, 7, 30, seed=3)
    cam = np.vstack([P.cam, P.cam[:1] + 0.1])
    pt = np.vstack([P.pt, P.pt[:2] + 0.1])
    S = capi.Solver(6, 9, P.cam_idx, P.pt_idx, P.obs)
    S.set_params(cam, pt)
    O = Oracle(6, 9, P.cam_idx, P.pt_idx, P.obs)
    cost = S.linearize()
    r, J = O.forward(cam, pt)
    Hpp, Hll, W, g = O.build(r, J)
    assert abs(cost - O.cost(r)) <= 1e-10 * O.cost(r)
    assert rel_fro(S.read(capi.BUF_HPP), Hpp) <= 1e-10 and np.all(S.read(capi.BUF_HPP)[5] == 0)
    assert rel_fro(S.read(capi.BUF_HLL), Hll) <= 1e-10 and np.all(S.read(capi.BUF_HLL)[7:] == 0)
    S.process_diag(1e3)
    it, done = S.solve(max_iter=10, tol=0.0, refuse_ratio=1e300)
    dx = S.read(capi.BUF_DX)
    assert np.all(np.isfinite(dx)) and np.all(dx[9 * 5:9 * 6] == 0) and np.all(dx[9 * 6 + 21:] == 0)


@pytest.mark.parametrize("dtype", DTYPES)
def test_device_pointer_and_async_entry_points(dtype):
    """ba_set_params_device / ba_get_params_device (caller-owned device buffers, ordered through the caller's stream) and
    ba_solve_async / ba_solve_wait give the results of their host-buffer / blocking counterparts."""
    import torch
    P = medium_problem()
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    cam_d = torch.tensor(P.cam, dtype=tdt, device="cuda")
    pt_d = torch.tensor(P.pt, dtype=tdt, device="cuda")
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    A = capi.Solver(P.nC, P.nP, P.cam_idx, P.pt_idx, P.obs, dtype=dtype)
    B = capi.Solver(P.nC, P.nP, P.cam_idx, P.pt_idx, P.obs, dtype=dtype)
    A.set_params(P.cam, P.pt)
    with torch.cuda.stream(st):
        cam_s, pt_s = cam_d * 1.0, pt_d * 1.0  # produced ON the caller's stream: the solver must wait for them
        B.set_params_device(cam_s.data_ptr(), pt_s.data_ptr(), st.cuda_stream)
    ca, cb = A.linearize(), B.linearize()
    assert ca == cb
    A.process_diag(1e3)
    B.process_diag(1e3)
    ita, _ = A.solve(max_iter=12, tol=0.0, refuse_ratio=1e300)
    B.solve_async(max_iter=12, tol=0.0, refuse_ratio=1e300)
    with pytest.raises(RuntimeError):
        B.solve_async(max_iter=12, tol=0.0, refuse_ratio=1e300)  # one pending solve at a time
    itb, _ = B.solve_wait()
    assert ita == itb == 12
    t = TOL[np.dtype(dtype)]["dx"]
    assert rel_fro(B.read(capi.BUF_DX), A.read(capi.BUF_DX)) <= t  # same kernels; only the RED order differs
    # early exit by the reference's rules without host polling
    B.solve_async(max_iter=100, tol=1e-1, refuse_ratio=1.0)
    itc, donec = B.solve_wait()
    itd, doned = A.solve(max_iter=100, tol=1e-1, refuse_ratio=1.0)  # same warm start (the 12-iteration step) as B had
    assert (itc, donec) == (itd, doned)
    A.accept_step()
    B.accept_step()
    cam_o, pt_o = torch.empty_like(cam_d), torch.empty_like(pt_d)
    with torch.cuda.stream(st):
        B.get_params_device(cam_o.data_ptr(), pt_o.data_ptr(), st.cuda_stream)
        cam_o2 = cam_o + 0.0  # consumed on the caller's stream without a host synchronisation
    st.synchronize()
    cam_b, pt_b = B.get_params()
    assert np.array_equal(cam_o2.cpu().numpy(), cam_b) and np.array_equal(pt_o.cpu().numpy(), pt_b)
    cam_a, pt_a = A.get_params()
    assert rel_fro(cam_b, cam_a) <= t and rel_fro(pt_b, pt_a) <= t

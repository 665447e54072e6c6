"""GPU parity of every (layout, operator stream) flavour of the hot path against the CPU oracle.

The default of ba_create is the ELL layout with the factored operator stream; the tile layout
with the W / J streams is the measured baseline, and the ELL layout can also stream the full Jacobian (which it does
by itself while any camera is in the small-angle branch).  test_gpu_parity.py runs everything on the default; this
file runs the linear-algebra core on every flavour so that none of them can rot."""
import numpy as np
import pytest

from megba_b200 import capi
from oracle.oracle import Oracle
from util import TOL, bal_like_problem, bigpoint_problem, medium_problem, rel_fro, small_problem

pytestmark = pytest.mark.gpu

# (layout, store, implicit)
FLAVOURS = {
    "ell-fact": (capi.LAYOUT_ELL, capi.STORE_FACT, False),
    "ell-fact-implicit-kind": (capi.LAYOUT_ELL, capi.STORE_FACT, True),
    "ell-j": (capi.LAYOUT_ELL, capi.STORE_J, False),
    "tile-w": (capi.LAYOUT_TILE, capi.STORE_AUTO, False),
    "tile-j": (capi.LAYOUT_TILE, capi.STORE_AUTO, True),
}
PROBLEMS = {"small": small_problem, "medium": medium_problem, "bigpoint": bigpoint_problem, "bal": bal_like_problem}


def _solver(P, dtype, flavour):
    layout, store, implicit = FLAVOURS[flavour]
    S = capi.Solver(P.nC, P.nP, P.cam_idx, P.pt_idx, P.obs, dtype=dtype, implicit=implicit, layout=layout, store=store)
    S.set_params(P.cam, P.pt)
    return S


def _oracle_step(P, dtype, region, n_it, info=None, dx0=None):
    O = Oracle(P.nC, P.nP, P.cam_idx, P.pt_idx, P.obs, dtype)
    r, J = O.forward(P.cam, P.pt)
    Hpp, Hll, W, g = O.build(r, J, info=info)
    HppD, _ = O.process_diag(Hpp, region)
    HllD, _ = O.process_diag(Hll, region)
    x0 = np.zeros(9 * P.nC + 3 * P.nP) if dx0 is None else dx0
    dx, it, _, _ = O.schur_pcg(HppD, HllD, W, None, g, x0, n_it, 0.0, 1e300)
    return O, (r, J, Hpp, Hll, W, g, HppD, HllD), dx, it


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("name", ["small", "medium", "bigpoint", "bal"])
@pytest.mark.parametrize("flavour", list(FLAVOURS))
def test_flavour_linearize_operator_pcg(flavour, name, dtype):
    P = PROBLEMS[name]()
    S = _solver(P, dtype, flavour)
    t = TOL[np.dtype(dtype)]
    region, n_it = 1e3, 10
    O, (r, J, Hpp, Hll, W, g, HppD, HllD), dx0, it0 = _oracle_step(P, dtype, region, n_it)
    cost = S.linearize()
    assert abs(cost - O.cost(r)) <= t["cost"] * O.cost(r)
    assert rel_fro(S.read(capi.BUF_HPP), Hpp) <= t["H"]
    assert rel_fro(S.read(capi.BUF_HLL), Hll) <= t["H"]
    assert rel_fro(S.read(capi.BUF_G), g) <= t["H"]
    S.process_diag(region)
    p = np.random.default_rng(3).normal(size=9 * P.nC)
    q = S.schur_apply(p)
    q0 = O.schur_mul(HppD, O.block_inverse(HllD), W, None, p)
    assert rel_fro(q, q0) <= 10 * t["H"]
    it, _ = S.solve(max_iter=n_it, tol=0.0, refuse_ratio=1e300)
    assert it == it0 == n_it
    assert rel_fro(S.read(capi.BUF_DX), dx0) <= t["dx"]
    # a second solve warm-starts from the first (schur_pcg_solver.cu:193,206) -- exercises the INIT pass on x0 != 0
    _, _, dx1, _ = _oracle_step(P, dtype, region, 4, dx0=dx0)
    S.solve(max_iter=4, tol=0.0, refuse_ratio=1e300)
    assert rel_fro(S.read(capi.BUF_DX), dx1) <= 10 * t["dx"]


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("flavour", ["ell-fact", "ell-fact-implicit-kind", "ell-j", "tile-w"])
def test_flavour_information_matrices(flavour, dtype):
    """H = J^T Omega J through every stream that supports it -- including the implicit compute kind of the ELL
    layout (upstream's implicit build silently ignores Omega, build_implicit_linear_system.cu:272-274)."""
    P = small_problem()
    rng = np.random.default_rng(11)
    a, b, c = rng.uniform(0.5, 2.0, P.nE), rng.uniform(-0.3, 0.3, P.nE), rng.uniform(0.5, 2.0, P.nE)
    info = np.stack([a, b, b, c], axis=1)
    S = _solver(P, dtype, flavour)
    S.set_information(info)
    t = TOL[np.dtype(dtype)]
    region, n_it = 1e3, 8
    O, (r, J, Hpp, Hll, W, g, HppD, HllD), dx0, it0 = _oracle_step(P, dtype, region, n_it, info=info)
    S.linearize()
    assert rel_fro(S.read(capi.BUF_HPP), Hpp) <= t["H"]
    assert rel_fro(S.read(capi.BUF_HLL), Hll) <= t["H"]
    assert rel_fro(S.read(capi.BUF_G), g) <= t["H"]
    S.process_diag(region)
    p = np.random.default_rng(5).normal(size=9 * P.nC)
    q0 = O.schur_mul(HppD, O.block_inverse(HllD), W, None, p)
    assert rel_fro(S.schur_apply(p), q0) <= 10 * t["H"]
    S.solve(max_iter=n_it, tol=0.0, refuse_ratio=1e300)
    assert rel_fro(S.read(capi.BUF_DX), dx0) <= t["dx"]


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_small_angle_camera_falls_back_to_the_jacobian_stream(dtype):
    """A camera with theta^2 <= eps takes the reference's first-order branch (analytical_derivatives.cu:131-158), where
    R = I + [omega]x is not a rotation; the factored stream then gives way to the Jacobian stream for that linearisation
    and returns once the camera has moved on."""
    P = small_problem()
    cam = P.cam.copy()
    cam[2, :3] = 0.0
    cam[5, :3] = [1e-9, -2e-9, 5e-10]
    Q = type(P)(P.nC, P.nP, P.nE, P.cam_idx, P.pt_idx, P.obs, cam, P.pt, "small-angle")
    S = _solver(Q, dtype, "ell-fact")
    t = TOL[np.dtype(dtype)]
    region, n_it = 1e3, 10
    O, sysm, dx0, _ = _oracle_step(Q, dtype, region, n_it)
    S.linearize()
    S.process_diag(region)
    S.solve(max_iter=n_it, tol=0.0, refuse_ratio=1e300)
    assert rel_fro(S.read(capi.BUF_DX), dx0) <= t["dx"]
    a_pcg_small, _ = S.algorithmic_bytes()
    # LM moves the cameras away from the degenerate branch: the trace must match the oracle's
    S.set_params(Q.cam, Q.pt)  # deltaX back to zero (the solve above would warm-start the first PCG)
    tr = S.lm_solve(lm_max_iter=4, initial_region=1e3, pcg_max_iter=20, pcg_tol=0.0, pcg_refuse=1e300, eps1=0.0, eps2=0.0)
    _, _, tr0 = O.lm_solve(Q.x(dtype), lm_max_iter=4, initial_region=1e3, eps1=0.0, eps2=0.0, pcg_max_iter=20,
                           pcg_tol=0.0, pcg_refuse=1e300)
    assert [r["accepted"] for r in tr] == [r["accepted"] for r in tr0]
    for a, b in zip(tr, tr0):
        assert abs(a["cost"] - b["cost"]) <= 50 * t["cost"] * abs(b["cost"])
    a_pcg_after, _ = S.algorithmic_bytes()
    assert a_pcg_after < a_pcg_small  # back on the 8-scalar stream


def test_layout_store_combinations_are_validated():
    P = small_problem()
    with pytest.raises(RuntimeError):
        capi.Solver(P.nC, P.nP, P.cam_idx, P.pt_idx, P.obs, layout=capi.LAYOUT_ELL, store=capi.STORE_W)
    with pytest.raises(RuntimeError):
        capi.Solver(P.nC, P.nP, P.cam_idx, P.pt_idx, P.obs, layout=capi.LAYOUT_TILE, store=capi.STORE_FACT)


@pytest.mark.parametrize("flavour", ["ell-fact", "tile-w"])
def test_lm_trace_on_bal_like_graph(flavour):
    """Whole LM runs (accept / reject sequence, costs, final parameters) on the heavy-tailed, camera-coherent graph."""
    P = bal_like_problem()
    S = _solver(P, np.float64, flavour)
    O = Oracle(P.nC, P.nP, P.cam_idx, P.pt_idx, P.obs, np.float64)
    kw = dict(lm_max_iter=6, initial_region=1e3, eps1=1e-3, eps2=1e-12, pcg_max_iter=30, pcg_tol=1e-3, pcg_refuse=1.0)
    tr = S.lm_solve(**kw)
    x1, _, tr0 = O.lm_solve(P.x(), **kw)
    assert [r["accepted"] for r in tr] == [r["accepted"] for r in tr0]
    assert [r["pcg_iters"] for r in tr] == [r["pcg_iters"] for r in tr0]
    assert np.allclose([r["cost"] for r in tr], [r["cost"] for r in tr0], rtol=1e-8)
    cam, pt = S.get_params()
    assert rel_fro(np.concatenate([cam.ravel(), pt.ravel()]), x1) <= 1e-7


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("name", ["medium", "bal"])
def test_persistent_pcg_kernel(name, dtype):
    """ba_set_persistent(1): the whole solve in one cooperative launch (edge phases with dynamic slice scheduling, camera
    phases spread over all SMs, grid barriers) gives the step of the two-launches-per-iteration path, including the
    reference's early exits."""
    P = PROBLEMS[name]()
    A, B = _solver(P, dtype, "ell-fact"), _solver(P, dtype, "ell-fact")
    B.set_persistent(True)
    t = TOL[np.dtype(dtype)]
    for S in (A, B):
        S.linearize()
        S.process_diag(1e3)
    ita = A.solve(max_iter=12, tol=0.0, refuse_ratio=1e300)
    itb = B.solve(max_iter=12, tol=0.0, refuse_ratio=1e300)
    assert ita == itb
    O, _, dx0, _ = _oracle_step(P, dtype, 1e3, 12)
    assert rel_fro(B.read(capi.BUF_DX), dx0) <= t["dx"]
    assert rel_fro(B.read(capi.BUF_DX), A.read(capi.BUF_DX)) <= t["dx"]
    ita = A.solve(max_iter=100, tol=1e-1, refuse_ratio=1.0)   # warm start + the README stopping rules
    itb = B.solve(max_iter=100, tol=1e-1, refuse_ratio=1.0)
    assert ita == itb
    assert rel_fro(B.read(capi.BUF_DX), A.read(capi.BUF_DX)) <= 10 * t["dx"]
    kw = dict(lm_max_iter=4, initial_region=1e3, eps1=1e-3, eps2=1e-12, pcg_max_iter=30, pcg_tol=1e-3, pcg_refuse=1.0)
    A.set_params(P.cam, P.pt)
    B.set_params(P.cam, P.pt)
    tra, trb = A.lm_solve(**kw), B.lm_solve(**kw)
    if dtype == np.float64:
        assert [r["accepted"] for r in tra] == [r["accepted"] for r in trb]
        assert [r["pcg_iters"] for r in tra] == [r["pcg_iters"] for r in trb]
        assert np.allclose([r["cost"] for r in tra], [r["cost"] for r in trb], rtol=50 * t["cost"])
    else:  # fp32: near convergence the accept / reject decision is within rounding of the summation order
        assert abs(tra[-1]["cost"] - trb[-1]["cost"]) <= 1e-3 * tra[-1]["cost"]


@pytest.mark.parametrize("split", ["1", "2", "4"])
@pytest.mark.parametrize("flavour,name", [("ell-fact", "bal"), ("ell-j", "bal"), ("ell-fact", "small"), ("ell-fact-implicit-kind", "medium")])
def test_split_slices(monkeypatch, flavour, name, split):
    """BA_ELL_SPLIT: 1, 2 or 4 warps share one slice of the S*p pass (the library picks 2 for shards of about one wave of
    slices, 1 otherwise); all three give the step of the oracle, with information matrices too."""
    monkeypatch.setenv("BA_ELL_SPLIT", split)
    P = PROBLEMS[name]()
    rng = np.random.default_rng(11)
    A = rng.normal(size=(P.nE, 2, 2))
    info = np.eye(2) + 0.2 * (A + A.transpose(0, 2, 1))
    t = TOL[np.dtype(np.float64)]
    for om in (None, info):
        S = _solver(P, np.float64, flavour)
        if om is not None:
            S.set_information(om)
        S.linearize()
        S.process_diag(1e3)
        it, _ = S.solve(max_iter=10, tol=0.0, refuse_ratio=1e300)
        _, _, dx0, it0 = _oracle_step(P, np.float64, 1e3, 10, info=om)
        assert it == it0
        assert rel_fro(S.read(capi.BUF_DX), dx0) <= t["dx"]
        cnew, msq = S.eval_trial()
        O = Oracle(P.nC, P.nP, P.cam_idx, P.pt_idx, P.obs)
        x1 = P.x() + dx0
        r1, _ = O.forward(x1[:9 * P.nC].reshape(-1, 9), x1[9 * P.nC:].reshape(-1, 3), jac=False)
        c1 = O.cost(r1) if om is None else None
        if c1 is not None:
            assert abs(cnew - c1) <= 1e-7 * c1
        S.close()

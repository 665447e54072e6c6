"""Pins the oracle (and the product) against the REFERENCE'S OWN KERNELS.

oracle/_ref/libmegba_ref.so is built by `make -C oracle/shim` from /root/reference/src/{geo/analytical_derivatives,
edge/build_linear_system,solver/schur_pcg_solver}.cu, compiled unmodified against shim headers (oracle/shim/README.md);
it is a git-ignored artefact that travels to the GPU box.  The reference has no CPU path, so this runs on the GPU only."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import ROOT
from megba_b200 import capi
from oracle.oracle import Oracle
from util import TOL, bigpoint_problem, medium_problem, rel_fro, rel_rowmax, small_problem

pytestmark = pytest.mark.gpu
REF = os.path.join(ROOT, "oracle", "_ref", "libmegba_ref.so")


@pytest.fixture(scope="module")
def ref():
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref/libmegba_ref.so not built (needs /root/reference at build time)")
    return C.CDLL(REF)


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _ref_forward(L, P, dtype):
    sfx = "_f64" if dtype == np.float64 else "_f32"
    r = np.empty((P.nE, 2), dtype)
    J = np.empty((P.nE, 2, 12), dtype)
    obs, cam, pt = (np.ascontiguousarray(a, dtype) for a in (P.obs, P.cam, P.pt))
    rc = getattr(L, "ref_forward" + sfx)(C.c_int64(P.nE), _p(P.cam_idx), _p(P.pt_idx), _p(obs), _p(cam), _p(pt), _p(r), _p(J))
    assert rc == 0
    return r, J


def _ref_build(L, P, r, J, info, dtype):
    sfx = "_f64" if dtype == np.float64 else "_f32"
    Hpp = np.empty((P.nC, 9, 9), dtype)
    Hll = np.empty((P.nP, 3, 3), dtype)
    W = np.empty((P.nE, 9, 3), dtype)
    g = np.empty(9 * P.nC + 3 * P.nP, dtype)
    info = None if info is None else np.ascontiguousarray(info, dtype)
    rc = getattr(L, "ref_build" + sfx)(C.c_int64(P.nC), C.c_int64(P.nP), C.c_int64(P.nE), _p(P.cam_idx), _p(P.pt_idx),
                                       _p(np.ascontiguousarray(r, dtype)), _p(np.ascontiguousarray(J, dtype)), _p(info),
                                       _p(Hpp), _p(Hll), _p(W), _p(g))
    assert rc == 0, rc  # -2 would mean Hlp is not the transpose of Hpl
    return Hpp, Hll, W, g


def _ref_solve(L, P, HppD, HllD, W, g, dx0, max_iter, tol, refuse, dtype):
    sfx = "_f64" if dtype == np.float64 else "_f32"
    dx = np.ascontiguousarray(dx0, dtype).copy()
    rc = getattr(L, "ref_solve" + sfx)(C.c_int64(P.nC), C.c_int64(P.nP), C.c_int64(P.nE), _p(P.cam_idx), _p(P.pt_idx),
                                       _p(np.ascontiguousarray(HppD, dtype)), _p(np.ascontiguousarray(HllD, dtype)),
                                       _p(np.ascontiguousarray(W, dtype)), _p(np.ascontiguousarray(g, dtype)), _p(dx),
                                       C.c_int(max_iter), C.c_double(tol), C.c_double(refuse))
    assert rc == 0
    return dx


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("name", ["small", "medium", "bigpoint"])
def test_a2_residual_jacobian_kernel(ref, name, dtype):
    """AnalyticalDerivativesKernelGradKernel (analytical_derivatives.cu:162-285) vs oracle vs product."""
    P = {"small": small_problem, "medium": medium_problem, "bigpoint": bigpoint_problem}[name]()
    if name == "bigpoint":
        P.cam[3, :3] = 0.0
        P.cam[4, :3] = [1e-9, -2e-9, 5e-10]
    r_ref, J_ref = _ref_forward(ref, P, dtype)
    O = Oracle(P.nC, P.nP, P.cam_idx, P.pt_idx, P.obs, dtype)
    r0, J0 = O.forward(P.cam, P.pt)
    t = TOL[np.dtype(dtype)]["rj"]
    assert np.abs(r0.astype(np.float64) - r_ref).max() <= 50 * t * max(1.0, np.abs(r_ref).max())
    assert rel_rowmax(J0, J_ref) <= 50 * t
    S = capi.Solver(P.nC, P.nP, P.cam_idx, P.pt_idx, P.obs, dtype=dtype)
    S.set_params(P.cam, P.pt)
    r1, J1 = S.eval_edges()
    assert np.abs(r1.astype(np.float64) - r_ref).max() <= 50 * t * max(1.0, np.abs(r_ref).max())
    assert rel_rowmax(J1, J_ref) <= 50 * t


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("with_info", [False, True])
def test_a3_assembly_kernels(ref, with_info, dtype):
    """makeHSchur / JMulInfo + makeHSchurWithInfo (build_linear_system.cu:88-239) vs oracle vs product."""
    P = medium_problem()
    O = Oracle(P.nC, P.nP, P.cam_idx, P.pt_idx, P.obs, dtype)
    r, J = O.forward(P.cam, P.pt)
    info = None
    if with_info:
        rng = np.random.default_rng(11)
        a, b, c = rng.uniform(0.5, 2.0, P.nE), rng.uniform(-0.3, 0.3, P.nE), rng.uniform(0.5, 2.0, P.nE)
        info = np.stack([a, b, b, c], axis=1)
    Hpp_r, Hll_r, W_r, g_r = _ref_build(ref, P, r, J, info, dtype)
    Hpp, Hll, W, g = O.build(r, J, info=info)
    t = TOL[np.dtype(dtype)]["H"]
    for got, want in ((Hpp, Hpp_r), (Hll, Hll_r), (W, W_r), (g, g_r)):
        assert rel_fro(got, want) <= t
    S = capi.Solver(P.nC, P.nP, P.cam_idx, P.pt_idx, P.obs, dtype=dtype)
    S.set_params(P.cam, P.pt)
    if with_info:
        S.set_information(info)
    S.linearize()
    for buf, want in ((capi.BUF_HPP, Hpp_r), (capi.BUF_HLL, Hll_r), (capi.BUF_W, W_r), (capi.BUF_G, g_r)):
        assert rel_fro(S.read(buf), want) <= t


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("rule", ["fixed12", "readme", "warm", "refuse"])
def test_a5_a8_schur_pcg_solver(ref, rule, dtype):
    """SchurPCGSolver::solve (schur_pcg_solver.cu:599-679: cublas matinvBatched, MakeV, cuSPARSE-SpMV PCG, SolveW)
    vs the oracle's restatement vs the product's ba_solve -- same damped blocks, same W, same g, same x0."""
    P = medium_problem()
    O = Oracle(P.nC, P.nP, P.cam_idx, P.pt_idx, P.obs, dtype)
    r, J = O.forward(P.cam, P.pt)
    Hpp, Hll, W, g = O.build(r, J)
    region = 1e4 if rule == "readme" else 1e3
    HppD, _ = O.process_diag(Hpp, region)
    HllD, _ = O.process_diag(Hll, region)
    nX = 9 * P.nC + 3 * P.nP
    x0 = np.zeros(nX)
    max_iter, tol, refuse = {"fixed12": (12, 0.0, 1e300), "readme": (100, 1e-1, 1.0), "warm": (5, 0.0, 1e300),
                            "refuse": (50, 0.0, 0.05)}[rule]
    if rule == "warm":
        x0, _, _, _ = O.schur_pcg(HppD, HllD, W, None, g, np.zeros(nX), 100, 1e-1, 1.0)
        x0 = x0 * 0.9
    dx_ref = _ref_solve(ref, P, HppD, HllD, W, g, x0, max_iter, tol, refuse, dtype)
    dx0, it0, done0, _ = O.schur_pcg(HppD, HllD, W, None, g, x0, max_iter, tol, refuse)
    t = TOL[np.dtype(dtype)]["dx"]
    assert rel_fro(dx0, dx_ref) <= t
    S = capi.Solver(P.nC, P.nP, P.cam_idx, P.pt_idx, P.obs, dtype=dtype)
    S.set_params(P.cam, P.pt)
    S.linearize()
    S.process_diag(region)
    S.write_dx(x0)
    it, done = S.solve(max_iter=max_iter, tol=tol, refuse_ratio=refuse)
    assert (it, done) == (it0, done0)
    assert rel_fro(S.read(capi.BUF_DX), dx_ref) <= t


def _ref_lm(L, P, dtype, log_path, **kw):
    sfx = "_f64" if dtype == np.float64 else "_f32"
    x = P.x(dtype).copy()
    obs = np.ascontiguousarray(P.obs, dtype)
    rc = getattr(L, "ref_lm_solve" + sfx)(C.c_int64(P.nC), C.c_int64(P.nP), C.c_int64(P.nE), _p(P.cam_idx), _p(P.pt_idx),
                                          _p(obs), _p(x), C.c_int(kw["lm_max_iter"]), C.c_double(kw["initial_region"]),
                                          C.c_double(kw["eps1"]), C.c_double(kw["eps2"]), C.c_int(kw["pcg_max_iter"]),
                                          C.c_double(kw["pcg_tol"]), C.c_double(kw["pcg_refuse"]), C.c_char_p(log_path.encode()))
    assert rc == 0
    import re
    costs, acc = [], []
    for line in open(log_path).read().splitlines():
        m = re.match(r"Start with error: (\S+),", line)
        if m:
            costs.append(2 * float(m.group(1)))
            acc.append(1)
        m = re.match(r"Iter (\d+) error: (\S+),", line)
        if m:
            costs.append(2 * float(m.group(2)))
            acc.append(1)
        if re.match(r"Iter (\d+) failed", line):
            costs.append(costs[-1])
            acc.append(0)
    return x, costs, acc


@pytest.mark.parametrize("dtype,hard", [(np.float64, True), (np.float64, False), (np.float32, False)])
def test_a4_a10_a11_reference_lm_loop(ref, tmp_path, dtype, hard):
    """The reference's WHOLE LM loop -- LMAlgo::solveCUDA (lm_algo.cu:139-223) with its scalar kernels (:26-136),
    processDiag / backup / rollback / applyUpdate (schur_LM_linear_system.cu:162-218), its analytical kernel, assembly
    kernels and cuSPARSE PCG solver, all compiled unmodified -- against the oracle's restatement and the product's
    ba_lm_solve: same accept/reject sequence, same printed costs (the reference prints 6 significant digits), same
    final parameter vector."""
    from megba_b200.synthetic import make_synthetic
    P = make_synthetic(30, 2000, 9000, seed=5)
    if hard:  # strongly perturbed start: the run contains genuinely rejected steps
        rng = np.random.default_rng(1)
        P.cam[:, :3] += rng.normal(0, 0.3, size=(P.nC, 3))
        P.cam[:, 3:6] += rng.normal(0, 2.0, size=(P.nC, 3))
        P.pt += rng.normal(0, 3.0, size=P.pt.shape)
    kw = dict(lm_max_iter=12 if hard else 5, initial_region=1e8 if hard else 1e4, eps1=1e-3, eps2=1e-12,
              pcg_max_iter=60, pcg_tol=1e-6, pcg_refuse=1.0)
    x_ref, costs_ref, acc_ref = _ref_lm(ref, P, dtype, str(tmp_path / "ref_lm.log"), **kw)
    assert len(costs_ref) >= 4
    if hard:
        assert 0 in acc_ref[:8] and 1 in acc_ref[2:]
    O = Oracle(P.nC, P.nP, P.cam_idx, P.pt_idx, P.obs, dtype)
    x0, k0, tr0 = O.lm_solve(P.x(), **kw)
    S = capi.Solver(P.nC, P.nP, P.cam_idx, P.pt_idx, P.obs, dtype=dtype)
    S.set_params(P.cam, P.pt)
    tr1 = S.lm_solve(**kw)
    cam1, pt1 = S.get_params()
    x1 = np.concatenate([cam1.ravel(), pt1.ravel()])
    n = len(costs_ref) if dtype == np.float64 else 3
    for name, tr in (("oracle", tr0), ("product", tr1)):
        assert [t["accepted"] for t in tr][:n] == acc_ref[:n], (name, tr, acc_ref)
        got = np.array([t["cost"] for t in tr][:n])
        assert np.allclose(got, costs_ref[:n], rtol=2e-5 if dtype == np.float64 else 2e-3), (name, got, costs_ref)
    if dtype == np.float64:
        assert rel_fro(x0, x_ref) <= 1e-7
        assert rel_fro(x1, x_ref) <= 1e-7


def _ref_build_implicit(L, P, r, J, dtype):
    sfx = "_f64" if dtype == np.float64 else "_f32"
    Hpp = np.empty((P.nC, 9, 9), dtype)
    Hll = np.empty((P.nP, 3, 3), dtype)
    g = np.empty(9 * P.nC + 3 * P.nP, dtype)
    rc = getattr(L, "ref_build_implicit" + sfx)(C.c_int64(P.nC), C.c_int64(P.nP), C.c_int64(P.nE), _p(P.cam_idx), _p(P.pt_idx),
                                                _p(np.ascontiguousarray(r, dtype)), _p(np.ascontiguousarray(J, dtype)),
                                                _p(Hpp), _p(Hll), _p(g))
    assert rc == 0, rc
    return Hpp, Hll, g


def _ref_solve_implicit(L, P, HppD, HllD, J, g, dx0, max_iter, tol, refuse, dtype):
    sfx = "_f64" if dtype == np.float64 else "_f32"
    dx = np.ascontiguousarray(dx0, dtype).copy()
    rc = getattr(L, "ref_solve_implicit" + sfx)(C.c_int64(P.nC), C.c_int64(P.nP), C.c_int64(P.nE), _p(P.cam_idx), _p(P.pt_idx),
                                                _p(np.ascontiguousarray(HppD, dtype)), _p(np.ascontiguousarray(HllD, dtype)),
                                                _p(np.ascontiguousarray(J, dtype)), _p(np.ascontiguousarray(g, dtype)), _p(dx),
                                                C.c_int(max_iter), C.c_double(tol), C.c_double(refuse))
    assert rc == 0
    return dx


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("name", ["small", "medium"])
def test_a3i_implicit_assembly_kernel(ref, name, dtype):
    """makeHppHllSchur (src/edge/build_implicit_linear_system.cu:65-111) through
    EdgeVector::buildImplicitLinearSystemCUDA (:208-322) vs the oracle's implicit build vs the product (matrix-free
    kind, both layouts)."""
    P = {"small": small_problem, "medium": medium_problem}[name]()
    O = Oracle(P.nC, P.nP, P.cam_idx, P.pt_idx, P.obs, dtype)
    r, J = O.forward(P.cam, P.pt)
    Hpp0, Hll0, _, g0 = O.build(r, J, explicit=False)
    Hpp, Hll, g = _ref_build_implicit(ref, P, r, J, dtype)
    t = TOL[np.dtype(dtype)]["H"]
    assert rel_fro(Hpp0, Hpp) <= t and rel_fro(Hll0, Hll) <= t and rel_fro(g0, g) <= t
    for layout in (capi.LAYOUT_TILE, capi.LAYOUT_ELL):
        S = capi.Solver(P.nC, P.nP, P.cam_idx, P.pt_idx, P.obs, dtype=dtype, implicit=True, layout=layout)
        S.set_params(P.cam, P.pt)
        S.linearize()
        assert rel_fro(S.read(capi.BUF_HPP), Hpp) <= t
        assert rel_fro(S.read(capi.BUF_HLL), Hll) <= t
        assert rel_fro(S.read(capi.BUF_G), g) <= t


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("rule", ["fixed12", "readme", "warm"])
def test_a7i_implicit_schur_pcg_solver(ref, rule, dtype):
    """ImplicitSchurPCGSolver::solve (src/solver/implicit_schur_pcg_solver.cu:661-720: inverses, MakeV :476-549, the
    matrix-free PCG :181-473 on implicitETMulx / implicitEMulx :21-90, SolveW :552-618) vs the oracle's matrix-free
    restatement vs the product.  The shim hands the reference zeroed pool blocks (its MakeV accumulates into an
    uncleared buffer, :500-528; see oracle/shim/README.md)."""
    P = medium_problem()
    O = Oracle(P.nC, P.nP, P.cam_idx, P.pt_idx, P.obs, dtype)
    r, J = O.forward(P.cam, P.pt)
    Hpp, Hll, _, g = O.build(r, J, explicit=False)
    region = 1e4 if rule == "readme" else 1e3
    HppD, _ = O.process_diag(Hpp, region)
    HllD, _ = O.process_diag(Hll, region)
    nX = 9 * P.nC + 3 * P.nP
    x0 = np.zeros(nX)
    max_iter, tol, refuse = {"fixed12": (12, 0.0, 1e300), "readme": (100, 1e-1, 1.0), "warm": (5, 0.0, 1e300)}[rule]
    if rule == "warm":
        x0, _, _, _ = O.schur_pcg(HppD, HllD, None, J, g, np.zeros(nX), 100, 1e-1, 1.0)
        x0 = x0 * 0.9
    dx_ref = _ref_solve_implicit(ref, P, HppD, HllD, J, g, x0, max_iter, tol, refuse, dtype)
    dx0, it0, done0, _ = O.schur_pcg(HppD, HllD, None, J, g, x0, max_iter, tol, refuse)
    t = TOL[np.dtype(dtype)]["dx"]
    assert rel_fro(dx0, dx_ref) <= t
    for layout in (capi.LAYOUT_TILE, capi.LAYOUT_ELL):
        S = capi.Solver(P.nC, P.nP, P.cam_idx, P.pt_idx, P.obs, dtype=dtype, implicit=True, layout=layout)
        S.set_params(P.cam, P.pt)
        S.linearize()
        S.process_diag(region)
        S.write_dx(x0)
        it, done = S.solve(max_iter=max_iter, tol=tol, refuse_ratio=refuse)
        assert (it, done) == (it0, done0)
        assert rel_fro(S.read(capi.BUF_DX), dx_ref) <= t

"""User-defined residuals on the GPU (ba_set_residual_program): the program the facade traces from the reference's
automatic-differentiation BAL edge, evaluated by the device interpreter, against the oracle and against the
hand-written analytical kernels; a residual that is NOT the BAL model against the numpy dual-number interpreter."""
import os
import re
import subprocess
import threading

import numpy as np
import pytest

from megba_b200 import capi
from megba_b200.synthetic import make_synthetic, read_bal, write_bal
from oracle.oracle import Oracle
from residual_program import BIN, Asm as _Asm, build_examples, edge_inputs, eval_program, random_program as _random_program, trace_bal_edge
from util import bal_like_problem, rel_fro

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["jit", "interpreter"])
def evaluation_mode(request, monkeypatch):
    """Every test runs twice: with the NVRTC-generated kernels (the default) and with the device interpreter."""
    monkeypatch.setenv("BA_PROG_JIT", "1" if request.param == "jit" else "0")
    return 2 if request.param == "jit" else 1


def _solver(P, dtype=np.float64, prog=None, **kw):
    S = capi.Solver(P.nC, P.nP, P.cam_idx, P.pt_idx, P.obs, dtype=dtype, **kw)
    assert S.residual_program_mode() == 0
    if prog is not None:
        S.set_residual_program(prog["code"], prog["consts"], prog["n_slots"], prog["out"])
        want = 2 if os.environ.get("BA_PROG_JIT") != "0" else 1
        assert S.residual_program_mode() == want, S.L.ba_last_error(S.h).decode()
    S.set_params(P.cam, P.pt)
    return S


@pytest.mark.parametrize("dtype,tol_r,tol_j", [(np.float64, 1e-13, 1e-11), (np.float32, 2e-5, 2e-4)])
def test_interpreter_matches_oracle_forward(tmp_path, dtype, tol_r, tol_j):
    prog = trace_bal_edge(tmp_path)
    P = make_synthetic(30, 2500, 12000, seed=41)
    P.cam[4, :3] = 0.
    P.cam[9, :3] *= 1e-9
    S = _solver(P, dtype, prog)
    r, J = S.eval_edges()
    O = Oracle(P.nC, P.nP, P.cam_idx, P.pt_idx, P.obs)
    r0, J0 = O.forward(P.cam, P.pt)
    assert np.isfinite(r).all() and np.isfinite(J).all()
    assert rel_fro(r, r0) <= tol_r and rel_fro(J, J0) <= tol_j
    S.close()


@pytest.mark.parametrize("info", [False, True])
def test_linearisation_and_lm_match_the_analytical_kernels(tmp_path, info):
    prog = trace_bal_edge(tmp_path)
    P = make_synthetic(40, 3000, 15000, seed=42)
    rng = np.random.default_rng(0)
    A = rng.normal(size=(P.nE, 2, 2))
    Om = (np.eye(2) + 0.2 * (A + A.transpose(0, 2, 1))) if info else None
    out = []
    for pr in (None, prog):
        S = _solver(P, prog=pr)
        if info:
            S.set_information(Om)
        cost = S.linearize()
        bufs = [S.read(w) for w in (capi.BUF_HPP, capi.BUF_HLL, capi.BUF_G)]
        S.process_diag(1e3)
        it, _ = S.solve(max_iter=20, tol=0.0, refuse_ratio=1e300)
        dx = S.read(capi.BUF_DX)
        cnew, msq = S.eval_trial()
        S.set_params(P.cam, P.pt)
        tr = S.lm_solve(lm_max_iter=6, initial_region=1e4, eps1=1e-3, eps2=1e-12, pcg_max_iter=40, pcg_tol=1e-6)
        cam, pt = S.get_params()
        out.append(dict(cost=cost, bufs=bufs, dx=dx, cnew=cnew, msq=msq, tr=tr, cam=cam, pt=pt))
        S.close()
    a, b = out
    assert abs(a["cost"] - b["cost"]) <= 1e-12 * a["cost"]
    for x, y in zip(a["bufs"], b["bufs"]):
        assert rel_fro(y, x) <= 1e-10
    assert rel_fro(b["dx"], a["dx"]) <= 1e-7
    assert abs(a["cnew"] - b["cnew"]) <= 1e-9 * a["cnew"] and abs(a["msq"] - b["msq"]) <= 1e-8 * abs(a["msq"])
    assert [t["accepted"] for t in a["tr"]] == [t["accepted"] for t in b["tr"]]
    assert np.allclose([t["cost"] for t in a["tr"]], [t["cost"] for t in b["tr"]], rtol=1e-8)
    assert rel_fro(b["cam"], a["cam"]) <= 1e-7 and rel_fro(b["pt"], a["pt"]) <= 1e-7
    assert b["tr"][-1]["cost"] < 0.5 * b["tr"][0]["cost"]


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_heavy_tailed_graph(tmp_path, dtype):
    """Points with a warp of their own (more than 32 observations) and slices of every height, both precisions."""
    prog = trace_bal_edge(tmp_path)
    P = bal_like_problem()
    tol = 1e-10 if dtype == np.float64 else 2e-4
    out = []
    for pr in (None, prog):
        S = _solver(P, dtype, pr)
        cost = S.linearize()
        bufs = [S.read(w) for w in (capi.BUF_HPP, capi.BUF_HLL, capi.BUF_G)]
        S.process_diag(1e3)
        S.solve(max_iter=8, tol=0.0, refuse_ratio=1e300)
        out.append((cost, bufs, S.read(capi.BUF_DX), S.eval_trial()))
        S.close()
    (ca, ba_, dxa, ta), (cb, bb, dxb, tb) = out
    assert abs(ca - cb) <= tol * ca
    for x, y in zip(ba_, bb):
        assert rel_fro(y, x) <= tol
    assert rel_fro(dxb, dxa) <= (1e-7 if dtype == np.float64 else 5e-3)
    assert abs(ta[0] - tb[0]) <= (1e-9 if dtype == np.float64 else 1e-3) * ta[0]


def test_program_can_be_cleared_and_fixed_vertices_apply(tmp_path):
    prog = trace_bal_edge(tmp_path)
    P = make_synthetic(25, 1500, 8000, seed=43)
    cf = np.zeros(P.nC, np.uint8); cf[[0, 5]] = 1
    pf = np.zeros(P.nP, np.uint8); pf[::7] = 1
    res = []
    for pr in (None, prog):
        S = _solver(P, prog=pr)
        S.set_fixed(cf, pf)
        tr = S.lm_solve(lm_max_iter=4, initial_region=1e4, eps1=1e-3, eps2=1e-12, pcg_max_iter=30, pcg_tol=1e-6)
        cam, pt = S.get_params()
        res.append((tr, cam, pt))
        if pr is not None:  # back to the analytical model on the same handle
            S.set_residual_program(None)
            S.set_params(P.cam, P.pt)
            tr2 = S.lm_solve(lm_max_iter=4, initial_region=1e4, eps1=1e-3, eps2=1e-12, pcg_max_iter=30, pcg_tol=1e-6)
            assert np.allclose([t["cost"] for t in tr2], [t["cost"] for t in res[0][0]], rtol=1e-12)
        S.close()
    (tra, cama, pta), (trb, camb, ptb) = res
    assert np.allclose([t["cost"] for t in tra], [t["cost"] for t in trb], rtol=1e-8)
    assert np.array_equal(camb[[0, 5]], P.cam[[0, 5]]) and np.array_equal(ptb[::7], P.pt[::7])
    assert rel_fro(camb, cama) <= 1e-7 and rel_fro(ptb, pta) <= 1e-7


def _pinhole_euler_program():
    """A residual that is not the BAL model: Euler rotations about z, x, y by camera[0:3], translation camera[3:6],
    pinhole focal length camera[6] and principal point camera[7:9], with a smooth sqrt-regularised depth."""
    A = _Asm()
    X = [A.inp(9 + k) for k in range(3)]
    cz, sz = A.cos(A.inp(0)), A.sin(A.inp(0))
    cx, sx = A.cos(A.inp(1)), A.sin(A.inp(1))
    cy, sy = A.cos(A.inp(2)), A.sin(A.inp(2))
    x1 = A.sub(A.mul(cz, X[0]), A.mul(sz, X[1]))
    y1 = A.add(A.mul(sz, X[0]), A.mul(cz, X[1]))
    y2 = A.sub(A.mul(cx, y1), A.mul(sx, X[2]))
    z2 = A.add(A.mul(sx, y1), A.mul(cx, X[2]))
    x3 = A.add(A.mul(cy, x1), A.mul(sy, z2))
    z3 = A.sub(A.mul(cy, z2), A.mul(sy, x1))
    px, py, pz = A.add(x3, A.inp(3)), A.add(y2, A.inp(4)), A.add(z3, A.inp(5))
    depth = A.add(A.sqrt(A.add(A.mul(pz, pz), A.const(1e-6))), A.mul(A.const(1e-9), A.abs(pz)))
    f = A.inp(6)
    r0 = A.sub(A.sub(A.inp(7), A.div(A.mul(f, px), depth)), A.inp(12))
    r1 = A.sub(A.sub(A.inp(8), A.div(A.mul(f, py), depth)), A.inp(13))
    return A.prog(r0, r1)


def test_a_residual_that_is_not_bal(tmp_path):
    prog = _pinhole_euler_program()
    assert prog["n_slots"] <= 48
    P = make_synthetic(20, 1200, 6000, seed=44)
    # measurements generated by the model itself at perturbed parameters, so that the problem has a zero-residual optimum
    rng = np.random.default_rng(3)
    r_true, _ = eval_program(prog, edge_inputs(P))
    P.obs = (P.obs + r_true).copy()          # residual is now exactly zero at (P.cam, P.pt)
    cam0, pt0 = P.cam.copy(), P.pt.copy()
    P.cam = cam0 + 1e-3 * rng.normal(size=cam0.shape) * np.abs(cam0).clip(1e-2)
    P.pt = pt0 + 1e-3 * rng.normal(size=pt0.shape)
    S = _solver(P, prog=prog)
    r, J = S.eval_edges()
    r0, J0 = eval_program(prog, edge_inputs(P))
    assert rel_fro(r, r0) <= 1e-12 and rel_fro(J, J0) <= 1e-11
    cost = S.linearize()
    assert abs(cost - (r0 ** 2).sum()) <= 1e-10 * cost or abs(cost - 0.5 * (r0 ** 2).sum()) <= 1e-10 * cost
    g = S.read(capi.BUF_G)
    gc = np.zeros((P.nC, 9)); gl = np.zeros((P.nP, 3))
    np.add.at(gc, P.cam_idx, -np.einsum("eri,er->ei", J0[:, :, :9], r0))
    np.add.at(gl, P.pt_idx, -np.einsum("eri,er->ei", J0[:, :, 9:], r0))
    assert rel_fro(g[:9 * P.nC], gc.ravel()) <= 1e-10
    tr = S.lm_solve(lm_max_iter=15, initial_region=1e3, eps1=1e-12, eps2=1e-14, pcg_max_iter=60, pcg_tol=1e-8)
    assert tr[-1]["cost"] < 1e-6 * tr[0]["cost"]  # Gauss-Newton on a zero-residual problem converges fast
    S.close()


@pytest.mark.parametrize("seed", [1, 2, 3, 4, 5, 6])
def test_random_programs_against_the_numpy_interpreter(seed):
    """Fuzz of the whole chain -- validation, camera / observation split, group masks, interpreter (ba_eval_edges) and the
    generated kernels or the interpreter inside ba_linearize -- on random programs with register reuse."""
    try:
        prog = _random_program(seed)
    except OverflowError:
        pytest.skip("more than 255 values")
    if prog["n_slots"] > 48:
        pytest.skip("needs more registers than the interpreter has")
    P = make_synthetic(15, 600, 3000, seed=100 + seed)
    S = _solver(P, prog=prog)
    r0, J0 = eval_program(prog, edge_inputs(P))
    assert np.isfinite(J0).all()
    r, J = S.eval_edges()
    assert rel_fro(r, r0) <= 1e-12 and rel_fro(J, J0) <= 1e-11
    cost = S.linearize()
    assert abs(cost - (r0 ** 2).sum()) <= 1e-11 * cost
    Hpp = np.zeros((P.nC, 9, 9)); Hll = np.zeros((P.nP, 3, 3)); gc = np.zeros((P.nC, 9)); gl = np.zeros((P.nP, 3))
    np.add.at(Hpp, P.cam_idx, np.einsum("eri,erj->eij", J0[:, :, :9], J0[:, :, :9]))
    np.add.at(Hll, P.pt_idx, np.einsum("eri,erj->eij", J0[:, :, 9:], J0[:, :, 9:]))
    np.add.at(gc, P.cam_idx, -np.einsum("eri,er->ei", J0[:, :, :9], r0))
    np.add.at(gl, P.pt_idx, -np.einsum("eri,er->ei", J0[:, :, 9:], r0))
    assert rel_fro(S.read(capi.BUF_HPP), Hpp) <= 1e-10
    assert rel_fro(S.read(capi.BUF_HLL), Hll) <= 1e-10
    assert rel_fro(S.read(capi.BUF_G), np.concatenate([gc.ravel(), gl.ravel()])) <= 1e-10
    S.close()


def test_rejects_malformed_programs(tmp_path):
    P = make_synthetic(5, 60, 200, seed=45)
    S = _solver(P)
    good = _pinhole_euler_program()
    bad_op = good["code"].copy(); bad_op[3, 0] = 12
    early = good["code"].copy(); early[0] = [2, 0, 5, 6, 0, 0]  # ADD of registers nothing wrote yet
    for code, consts, n_slots, out, what in (
            (bad_op, good["consts"], good["n_slots"], good["out"], "unknown op"),
            (early, good["consts"], good["n_slots"], good["out"], "before it is written"),
            (good["code"], good["consts"], 49, good["out"], "registers"),
            (good["code"], good["consts"][:0], good["n_slots"], good["out"], "constant"),
            (good["code"], good["consts"], good["n_slots"], (200, 0), "result")):
        with pytest.raises(RuntimeError, match=what):
            S.set_residual_program(code, consts, n_slots, out)
    S.close()
    T = capi.Solver(P.nC, P.nP, P.cam_idx, P.pt_idx, P.obs, layout=capi.LAYOUT_TILE)
    with pytest.raises(RuntimeError, match="ELL"):
        T.set_residual_program(good["code"], good["consts"], good["n_slots"], good["out"])
    T.close()


def test_two_ranks_on_one_device(tmp_path):
    prog = trace_bal_edge(tmp_path)
    P = make_synthetic(30, 2000, 10000, seed=46)
    ref = _solver(P)
    tr0 = ref.lm_solve(lm_max_iter=4, initial_region=1e4, eps1=1e-3, eps2=1e-12, pcg_max_iter=30, pcg_tol=1e-6)
    ref.close()
    ws = 2
    solvers = [capi.Solver(P.nC, P.nP, P.cam_idx, P.pt_idx, P.obs, device=0, rank=r, world_size=ws) for r in range(ws)]
    capi.Solver.comm_init_local(solvers)
    out, err = [None] * ws, [None] * ws

    def work(r):
        try:
            S = solvers[r]
            S.set_residual_program(prog["code"], prog["consts"], prog["n_slots"], prog["out"])
            S.set_params(P.cam, P.pt)
            out[r] = S.lm_solve(lm_max_iter=4, initial_region=1e4, eps1=1e-3, eps2=1e-12, pcg_max_iter=30, pcg_tol=1e-6)
        except Exception as exc:
            err[r] = exc

    th = [threading.Thread(target=work, args=(r,)) for r in range(ws)]
    [t.start() for t in th]
    [t.join(timeout=300) for t in th]
    assert not any(t.is_alive() for t in th)
    for e in err:
        if e is not None:
            raise e
    for S in solvers:
        S.close()
    assert np.allclose([t["cost"] for t in out[0]], [t["cost"] for t in tr0], rtol=1e-8)


@pytest.mark.parametrize("exe,dtype,implicit", [("BAL_Double", np.float64, False), ("BAL_Double_implicit", np.float64, True),
                                                ("BAL_Float", np.float32, False)])
def test_reference_autodiff_examples_run_through_the_facade(tmp_path, exe, dtype, implicit):
    """The reference's three automatic-differentiation drivers (examples/BAL_Double.cpp, BAL_Double_implicit.cpp,
    BAL_Float.cpp) against the oracle on the same file."""
    build_examples()
    P = make_synthetic(20, 900, 4000, seed=13)
    path = str(tmp_path / "problem-20-900-pre.txt")
    write_bal(path, P)
    flags = ["--max_iter", "4", "--solver_max_iter", "40", "--solver_tol", "1e-6", "--tau", "1e4", "--epsilon1", "1e-3",
             "--epsilon2", "1e-12"]
    r = subprocess.run([os.path.join(BIN, exe), "--path", path, *flags], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    m = re.search(r"RESULT iterations (\d+) final_cost (\S+) camera_sum (\S+)", r.stdout)
    assert m, r.stdout
    Q = read_bal(path)
    O = Oracle(Q.nC, Q.nP, Q.cam_idx, Q.pt_idx, Q.obs, dtype)
    x, k, tr = O.lm_solve(Q.x(), implicit=implicit, lm_max_iter=4, initial_region=1e4, eps1=1e-3, eps2=1e-12, pcg_max_iter=40,
                          pcg_tol=1e-6, pcg_refuse=1.0)
    tol = 1e-8 if dtype == np.float64 else 2e-3
    assert int(m.group(1)) == len(tr) - 1
    assert abs(float(m.group(2)) - tr[-1]["cost"]) <= tol * tr[-1]["cost"]
    assert abs(float(m.group(3)) - x[:9 * Q.nC].sum()) <= max(tol, 1e-9) * abs(x[:9 * Q.nC].sum()) * 10

"""Sharded solve with MORE RANKS THAN DEVICES: 2-4 point-sharded solvers on ONE GPU, driven from one host thread each
through the in-process communicator (ba_comm_init_local).  Exercises on a single-GPU box everything the multi-GPU path
is made of -- ba_partition_points, the per-rank ELL layouts, the fused push exchange inside the camera-space PCG
kernel (over plain device pointers here, over NVLink peer memory between GPUs), the once-per-linearisation
reductions and the point gather of ba_get_params -- against the single-rank oracle."""
import threading

import numpy as np
import pytest

from megba_b200 import capi
from megba_b200.synthetic import make_synthetic
from oracle.oracle import Oracle
from util import rel_fro

pytestmark = pytest.mark.gpu


def _run_ranks(P, ws, implicit, layout=capi.LAYOUT_AUTO):
    solvers = [capi.Solver(P.nC, P.nP, P.cam_idx, P.pt_idx, P.obs, dtype=np.float64, implicit=implicit, device=0, rank=r,
                           world_size=ws, layout=layout) for r in range(ws)]
    capi.Solver.comm_init_local(solvers)
    out = [None] * ws
    err = [None] * ws

    def work(r):
        try:
            S = solvers[r]
            S.set_params(P.cam, P.pt)
            cost = S.linearize()
            Hpp = S.read(capi.BUF_HPP)
            g = S.read(capi.BUF_G)
            S.process_diag(1e3)
            it, done = S.solve(max_iter=15, tol=0.0, refuse_ratio=1e300)
            dx = S.read(capi.BUF_DX)
            dxn, xn = S.step_norms()
            cnew, msq = S.eval_trial()
            S.set_params(P.cam, P.pt)
            tr = S.lm_solve(lm_max_iter=5, initial_region=1e4, eps1=1e-3, eps2=1e-12, pcg_max_iter=40, pcg_tol=1e-6)
            cam, pt = S.get_params()
            out[r] = dict(cost=cost, Hpp=Hpp, g=g, dx=dx, it=it, dxn=dxn, cnew=cnew, tr=tr, cam=cam, pt=pt)
        except Exception as exc:  # surfaced by the main thread
            err[r] = exc

    th = [threading.Thread(target=work, args=(r,)) for r in range(ws)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=300)
    assert not any(t.is_alive() for t in th), "a rank hung"
    for e in err:
        if e is not None:
            raise e
    for S in solvers:
        S.close()
    return out


@pytest.mark.parametrize("ws,implicit", [(2, False), (4, False), (3, True)])
@pytest.mark.timeout(600)
def test_ranks_sharing_one_gpu_match_single_rank_oracle(ws, implicit):
    P = make_synthetic(40, 3000, 14000, seed=21)
    R = _run_ranks(P, ws, implicit)
    O = Oracle(P.nC, P.nP, P.cam_idx, P.pt_idx, P.obs)
    r, J = O.forward(P.cam, P.pt)
    Hpp, Hll, W, g = O.build(r, J, explicit=not implicit)
    nc9 = 9 * P.nC
    for k in range(ws):
        assert abs(R[k]["cost"] - O.cost(r)) <= 1e-10 * O.cost(r)
        assert rel_fro(R[k]["Hpp"], Hpp) <= 1e-10 and np.array_equal(R[k]["Hpp"], R[0]["Hpp"])
        assert rel_fro(R[k]["g"][:nc9], g[:nc9]) <= 1e-10
    # point parts are rank-local: the ranks' point gradients tile the full vector
    assert rel_fro(sum(R[k]["g"][nc9:] for k in range(ws)), g[nc9:]) <= 1e-10
    HppD, _ = O.process_diag(Hpp, 1e3)
    HllD, _ = O.process_diag(Hll, 1e3)
    dx0, it0, _, _ = O.schur_pcg(HppD, HllD, W, J if implicit else None, g, np.zeros(nc9 + 3 * P.nP), 15, 0.0, 1e300)
    for k in range(ws):
        assert R[k]["it"] == it0 == 15
        assert rel_fro(R[k]["dx"][:nc9], dx0[:nc9]) <= 1e-8
        assert np.array_equal(R[k]["dx"][:nc9], R[0]["dx"][:nc9])  # the replicated camera step is bit-identical
        assert R[k]["dxn"] == R[0]["dxn"] and R[k]["cnew"] == R[0]["cnew"]
    assert rel_fro(sum(R[k]["dx"][nc9:] for k in range(ws)), dx0[nc9:]) <= 1e-8
    assert abs(R[0]["dxn"] - np.linalg.norm(dx0)) <= 1e-8 * np.linalg.norm(dx0)
    x1, _, tr = O.lm_solve(P.x(), implicit=implicit, lm_max_iter=5, initial_region=1e4, eps1=1e-3, eps2=1e-12,
                           pcg_max_iter=40, pcg_tol=1e-6, pcg_refuse=1.0)
    n = len(tr)
    for k in range(ws):
        assert [t["accepted"] for t in R[k]["tr"][:n]] == [t["accepted"] for t in tr]
        assert np.allclose([t["cost"] for t in R[k]["tr"][:n]], [t["cost"] for t in tr], rtol=1e-8)
        assert np.array_equal(R[k]["pt"], R[0]["pt"])  # ba_get_params gathers the full point array on every rank
    x_gpu = np.concatenate([R[0]["cam"].ravel(), R[0]["pt"].ravel()])
    assert rel_fro(x_gpu, x1) <= 1e-7


@pytest.mark.timeout(600)
def test_ranks_sharing_one_gpu_tile_layout():
    """The tile layout (W stream) through the same exchange."""
    P = make_synthetic(40, 3000, 14000, seed=22)
    R = _run_ranks(P, 2, False, layout=capi.LAYOUT_TILE)
    O = Oracle(P.nC, P.nP, P.cam_idx, P.pt_idx, P.obs)
    r, J = O.forward(P.cam, P.pt)
    Hpp, Hll, W, g = O.build(r, J)
    HppD, _ = O.process_diag(Hpp, 1e3)
    HllD, _ = O.process_diag(Hll, 1e3)
    nc9 = 9 * P.nC
    dx0, _, _, _ = O.schur_pcg(HppD, HllD, W, None, g, np.zeros(nc9 + 3 * P.nP), 15, 0.0, 1e300)
    assert rel_fro(R[0]["dx"][:nc9], dx0[:nc9]) <= 1e-8 and np.array_equal(R[0]["dx"][:nc9], R[1]["dx"][:nc9])
    assert rel_fro(R[0]["dx"][nc9:] + R[1]["dx"][nc9:], dx0[nc9:]) <= 1e-8

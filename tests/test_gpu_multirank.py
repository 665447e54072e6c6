"""Two ranks, two GPUs, NCCL + CUDA IPC: the point-sharded CUDA path reproduces the single-rank oracle.
Deselected when fewer than two GPUs are visible (tests/conftest.py; run with `gpurun --gpus 2`) -- the same sharded path with
all ranks on one device is tests/test_gpu_sharded_one_device.py."""
import os
import socket
import time

import numpy as np
import pytest

from conftest import HAVE_GPU
from oracle.oracle import Oracle
from util import rel_fro

pytestmark = pytest.mark.gpu


def _ngpu():
    if not HAVE_GPU:
        return 0
    import torch
    return torch.cuda.device_count()


def _worker(rank, ws, uid_path, out_path, implicit):
    import torch  # noqa: F401  (initialises CUDA libraries in the child)
    from megba_b200 import capi
    from megba_b200.synthetic import make_synthetic
    P = make_synthetic(40, 3000, 14000, seed=21)
    S = capi.Solver(P.nC, P.nP, P.cam_idx, P.pt_idx, P.obs, dtype=np.float64, implicit=implicit, device=rank, rank=rank,
                    world_size=ws)
    if rank == 0:
        with open(uid_path + ".tmp", "wb") as fh:
            fh.write(capi.Solver.comm_unique_id())
        os.rename(uid_path + ".tmp", uid_path)
    while not os.path.exists(uid_path):
        time.sleep(0.05)
    with open(uid_path, "rb") as fh:
        uid = fh.read()
    S.comm_init(uid)
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
    np.savez(out_path % rank, cost=cost, Hpp=Hpp, gc=g[:9 * P.nC], g=g, dx=dx, it=it, dxn=dxn, xn=xn, cnew=cnew, msq=msq,
             costs=np.array([t["cost"] for t in tr]), acc=np.array([t["accepted"] for t in tr]),
             pit=np.array([t["pcg_iters"] for t in tr]), cam=cam, pt=pt)
    S.close()


@pytest.mark.two_gpus
@pytest.mark.parametrize("implicit", [False, True])
@pytest.mark.timeout(600)
def test_two_gpus_match_single_rank_oracle(tmp_path, implicit):
    import torch.multiprocessing as mp
    from megba_b200.synthetic import make_synthetic
    uid_path = str(tmp_path / "uid.bin")
    out = str(tmp_path / "rank%d.npz")
    mp.spawn(_worker, args=(2, uid_path, out, implicit), nprocs=2, join=True)
    r0, r1 = np.load(out % 0), np.load(out % 1)
    P = make_synthetic(40, 3000, 14000, seed=21)
    O = Oracle(P.nC, P.nP, P.cam_idx, P.pt_idx, P.obs)
    r, J = O.forward(P.cam, P.pt)
    Hpp, Hll, W, g = O.build(r, J, explicit=not implicit)
    assert abs(r0["cost"] - O.cost(r)) <= 1e-10 * O.cost(r)
    assert rel_fro(r0["Hpp"], Hpp) <= 1e-10 and np.array_equal(r0["Hpp"], r1["Hpp"])
    assert rel_fro(r0["gc"], g[:9 * P.nC]) <= 1e-10
    # point parts are rank-local: the two ranks' point gradients tile the full vector
    assert rel_fro(r0["g"][9 * P.nC:] + r1["g"][9 * P.nC:], g[9 * P.nC:]) <= 1e-10
    HppD, _ = O.process_diag(Hpp, 1e3)
    HllD, _ = O.process_diag(Hll, 1e3)
    dx0, it0, _, _ = O.schur_pcg(HppD, HllD, W, J if implicit else None, g, np.zeros(9 * P.nC + 3 * P.nP), 15, 0.0, 1e300)
    assert int(r0["it"]) == it0 == 15
    nc9 = 9 * P.nC
    assert rel_fro(r0["dx"][:nc9], dx0[:nc9]) <= 1e-8
    assert np.array_equal(r0["dx"][:nc9], r1["dx"][:nc9])          # replicated camera step is bit-identical
    assert rel_fro(r0["dx"][nc9:] + r1["dx"][nc9:], dx0[nc9:]) <= 1e-8
    assert abs(r0["dxn"] - np.linalg.norm(dx0)) <= 1e-8 * np.linalg.norm(dx0)
    assert r0["dxn"] == r1["dxn"] and r0["cnew"] == r1["cnew"]
    x1, k, tr = O.lm_solve(P.x(), implicit=implicit, lm_max_iter=5, initial_region=1e4, eps1=1e-3, eps2=1e-12,
                           pcg_max_iter=40, pcg_tol=1e-6, pcg_refuse=1.0)
    n = len(tr)
    assert list(r0["acc"][:n]) == [t["accepted"] for t in tr]
    assert np.allclose(r0["costs"][:n], [t["cost"] for t in tr], rtol=1e-8)
    x_gpu = np.concatenate([r0["cam"].ravel(), r0["pt"].ravel()])
    assert rel_fro(x_gpu, x1) <= 1e-7
    assert np.array_equal(r0["pt"], r1["pt"])                       # get_params gathers the full point array on every rank


@pytest.mark.two_gpus
@pytest.mark.timeout(600)
def test_facade_world_size_two_threads_in_one_process(tmp_path):
    """The C++ facade drives deviceUsed = {0, 1} from two host threads of ONE process (like the reference's single
    process): NCCL ranks + the peer exchange through plain peer access instead of CUDA IPC."""
    import re
    import subprocess
    from conftest import ROOT
    from megba_b200.synthetic import make_synthetic, read_bal, write_bal
    env = dict(os.environ)
    env.pop("CC", None)
    subprocess.run(["make", "-C", os.path.join(ROOT, "examples"), "-s"], check=True, env=env)
    P = make_synthetic(20, 900, 4000, seed=13)
    path = str(tmp_path / "problem-20-900-pre.txt")
    write_bal(path, P)
    r = subprocess.run([os.path.join(ROOT, "build", "examples", "BAL_Double_analytical"), "--path", path, "--world_size", "2",
                        "--max_iter", "4", "--solver_max_iter", "40", "--solver_tol", "1e-6", "--tau", "1e4",
                        "--epsilon1", "1e-3", "--epsilon2", "1e-12"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    m = re.search(r"RESULT iterations (\d+) final_cost (\S+) camera_sum (\S+)", r.stdout)
    assert m, r.stdout
    Q = read_bal(path)
    O = Oracle(Q.nC, Q.nP, Q.cam_idx, Q.pt_idx, Q.obs)
    x, k, tr = O.lm_solve(Q.x(), lm_max_iter=4, initial_region=1e4, eps1=1e-3, eps2=1e-12, pcg_max_iter=40, pcg_tol=1e-6,
                          pcg_refuse=1.0)
    assert int(m.group(1)) == len(tr) - 1
    assert abs(float(m.group(2)) - tr[-1]["cost"]) <= 1e-8 * tr[-1]["cost"]
    assert abs(float(m.group(3)) - x[:9 * Q.nC].sum()) <= 1e-7 * abs(x[:9 * Q.nC].sum())

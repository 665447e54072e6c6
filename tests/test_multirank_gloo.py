"""world_size-2 host logic on CPU (gloo): the point partition computed by the C-ABI's
ba_partition_points shards the edges so that per-rank partial sums of Hpp / g_c / cost and of
the Schur operator's camera-side scatter, allreduced, equal the single-rank result; point-side
quantities are rank-local.  This is exactly the exchange pattern the CUDA path uses over NCCL
(DESIGN.md section 'multi-GPU'); the arithmetic here is the oracle's."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from megba_b200 import capi
from oracle.oracle import Oracle
from util import rel_fro, small_problem


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, ws, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        P = small_problem()
        deg = np.bincount(P.pt_idx, minlength=P.nP)
        cuts = capi.partition_points(deg, ws)
        mine = (P.pt_idx >= cuts[rank]) & (P.pt_idx < cuts[rank + 1])
        O = Oracle(P.nC, P.nP, P.cam_idx[mine], P.pt_idx[mine], P.obs[mine])
        r, J = O.forward(P.cam, P.pt)
        Hpp, Hll, W, g = O.build(r, J)
        cost = torch.tensor([O.cost(r)], dtype=torch.float64)
        tH = torch.from_numpy(np.concatenate([Hpp.ravel(), g[:9 * P.nC]]))
        dist.all_reduce(tH)
        dist.all_reduce(cost)
        Hpp_all = tH[:81 * P.nC].numpy().reshape(P.nC, 9, 9)
        # Schur operator: t and Hll^-1 are rank-local; only the 9nC scatter is reduced
        HppD, _ = O.process_diag(Hpp_all, 1e3)
        HllD, _ = O.process_diag(Hll, 1e3)
        own = slice(cuts[rank], cuts[rank + 1])
        HllInv = np.zeros_like(HllD)
        HllInv[own] = O.block_inverse(HllD[own])
        p = np.random.default_rng(5).normal(size=9 * P.nC)
        t = O.hlp_mul(W, None, p).reshape(P.nP, 3)
        u = np.einsum("pij,pj->pi", HllInv, t)
        s = torch.from_numpy(O.hpl_mul(W, None, u.ravel()))
        dist.all_reduce(s)
        q = np.einsum("cij,cj->ci", HppD, p.reshape(P.nC, 9)).ravel() - s.numpy()
        if rank == 0:
            np.savez(out, Hpp=Hpp_all, gc=tH[81 * P.nC:].numpy(), cost=cost.numpy(), q=q, cuts=cuts)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_partition_reproduces_single_rank(tmp_path):
    out = str(tmp_path / "r0.npz")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    P = small_problem()
    O = Oracle(P.nC, P.nP, P.cam_idx, P.pt_idx, P.obs)
    r, J = O.forward(P.cam, P.pt)
    Hpp, Hll, W, g = O.build(r, J)
    assert rel_fro(got["Hpp"], Hpp) < 1e-13
    assert rel_fro(got["gc"], g[:9 * P.nC]) < 1e-13
    assert abs(got["cost"][0] - O.cost(r)) < 1e-12 * O.cost(r)
    HppD, _ = O.process_diag(Hpp, 1e3)
    HllD, _ = O.process_diag(Hll, 1e3)
    p = np.random.default_rng(5).normal(size=9 * P.nC)
    q = O.schur_mul(HppD, O.block_inverse(HllD), W, None, p)
    assert rel_fro(got["q"], q) < 1e-12
    assert 0 < got["cuts"][1] < P.nP

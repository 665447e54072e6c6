"""Linearisation / LM-iteration time of the user-defined-residual path (device interpreter) beside the analytical
kernels, on the Venice-shaped problem.  Usage: python scripts/bench_program.py [workload]"""
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

from megba_b200 import capi
from megba_b200.synthetic import make_named
from residual_program import trace_bal_edge

P = make_named(sys.argv[1] if len(sys.argv) > 1 else "venice-1778")
prog = trace_bal_edge(tempfile.mkdtemp())
print(f"program: {len(prog['code'])} instructions, {prog['n_slots']} registers", flush=True)
_warm = capi.Solver(P.nC, P.nP, P.cam_idx, P.pt_idx, P.obs)  # context, module loading, allocator: not part of any line below
_warm.set_params(P.cam, P.pt)
_warm.lm_solve(lm_max_iter=2, initial_region=1e4, eps1=-1, eps2=-1, pcg_max_iter=5, pcg_tol=1e-2)
_warm.close()
for dtype in (np.float64, np.float32):
    for name, pr in (("analytical", None), ("program", prog)):
        S = capi.Solver(P.nC, P.nP, P.cam_idx, P.pt_idx, P.obs, dtype=dtype)
        if pr:
            S.set_residual_program(pr["code"], pr["consts"], pr["n_slots"], pr["out"])
        S.set_params(P.cam, P.pt)
        for _ in range(3):
            c = S.linearize()
        S.set_timing(True)
        S.reset_stats()
        for _ in range(10):
            S.linearize()
        st = S.stats()
        S.process_diag(1e4)
        S.solve(max_iter=30, tol=0.0, refuse_ratio=1e300)
        t0 = time.perf_counter()
        for _ in range(5):
            S.eval_trial()
        tt = (time.perf_counter() - t0) / 5
        S.set_timing(False)
        S.set_params(P.cam, P.pt)
        t0 = time.perf_counter()
        tr = S.lm_solve(lm_max_iter=10, initial_region=1e4, eps1=-1, eps2=-1, pcg_max_iter=30, pcg_tol=1e-2)
        dt = time.perf_counter() - t0
        name = name + {0: "", 1: "/interp", 2: "/jit"}[S.residual_program_mode()]
        print(f"{np.dtype(dtype).name:8s} {name:14s} linearise {st['linearize_ms']/st['linearize_calls']:8.3f} ms   trial {tt*1e3:7.3f} ms   "
              f"cost {c:.6e}   LM {len(tr)-1} it in {dt*1e3:7.1f} ms  final {tr[-1]['cost']:.6e}", flush=True)
        S.close()

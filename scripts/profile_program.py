import sys, tempfile
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
from megba_b200 import capi
from megba_b200.synthetic import make_named
from residual_program import trace_bal_edge
P = make_named("venice-1778")
prog = trace_bal_edge(tempfile.mkdtemp())
S = capi.Solver(P.nC, P.nP, P.cam_idx, P.pt_idx, P.obs)
S.set_residual_program(prog["code"], prog["consts"], prog["n_slots"], prog["out"])
S.set_params(P.cam, P.pt)
for _ in range(3):
    S.linearize()
S.process_diag(1e4); S.solve(max_iter=3, tol=0.0, refuse_ratio=1e300)
S.eval_trial()

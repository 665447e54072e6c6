"""Short, small-problem pass over every kernel flavour for `compute-sanitizer` (memcheck / racecheck / synccheck):
   compute-sanitizer --tool racecheck python scripts/sanitize_target.py
Problems: tests/util.py's `small` (ragged tiles), `bigpoint` (a point spanning several slabs + its empty tiles) and
a 3000-camera graph (cooperative-grid flavour of the camera kernel)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

from megba_b200 import capi  # noqa: E402
from megba_b200.synthetic import make_synthetic  # noqa: E402
from util import bigpoint_problem, small_problem  # noqa: E402

problems = [small_problem(), bigpoint_problem()]
if "--grid" in sys.argv:
    problems.append(make_synthetic(3000, 8000, 60000, seed=17))
for P in problems:
    for dtype in (np.float64, np.float32):
        for implicit in (False, True):
            S = capi.Solver(P.nC, P.nP, P.cam_idx, P.pt_idx, P.obs, dtype=dtype, implicit=implicit)
            S.set_params(P.cam, P.pt)
            c0 = S.linearize()
            S.process_diag(1e3)
            it, _ = S.solve(max_iter=4, tol=0.0, refuse_ratio=1e300)
            S.step_norms()
            S.eval_trial()
            tr = S.lm_solve(lm_max_iter=2, initial_region=1e3, eps1=0.0, eps2=0.0, pcg_max_iter=3, pcg_tol=0.0, pcg_refuse=1e300)
            print(P.name if hasattr(P, "name") else "?", np.dtype(dtype).name, "implicit" if implicit else "explicit",
                  f"cost {c0:.6g} -> {tr[-1]['cost']:.6g}, pcg {it}", flush=True)
            del S
print("sanitize target done")

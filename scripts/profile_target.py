"""Short workload for ncu: Venice-shaped problem, one linearisation and a few PCG iterations."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from megba_b200 import capi  # noqa: E402
from megba_b200.synthetic import make_named  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="venice-1778")
ap.add_argument("--pcg", type=int, default=4)
ap.add_argument("--implicit", action="store_true")
ap.add_argument("--dtype", default="f64")
a = ap.parse_args()
if os.environ.get("BA_LIB"):  # an experimental build (see scripts/bench_variants.py)
    capi.LIB_PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.environ["BA_LIB"])
P = make_named(a.workload)
S = capi.Solver(P.nC, P.nP, P.cam_idx, P.pt_idx, P.obs, dtype=np.float64 if a.dtype == "f64" else np.float32,
                implicit=a.implicit)
S.set_params(P.cam, P.pt)
print("cost", S.linearize())
S.process_diag(1e4)
print("solve", S.solve(max_iter=a.pcg, tol=0.0, refuse_ratio=1e300))
print("norms", S.step_norms(), "trial", S.eval_trial())

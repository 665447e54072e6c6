"""Generates tests/golden/oracle_lm_small.json from the CPU oracle.

Upstream MegBA ships no golden vectors; this fixture pins OUR oracle against accidental
drift (it is a regression fixture, not a reference-derived one).  Run from the repo root:
    python tests/golden/make_golden.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from megba_b200.synthetic import make_synthetic  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402

shape, seed = (12, 400, 1700), 1
params = dict(lm_max_iter=6, initial_region=1e3, eps1=1.0, eps2=1e-10, pcg_max_iter=30, pcg_tol=1e-1, pcg_refuse=1.0)
P = make_synthetic(*shape, seed=seed)
_, _, tr = Oracle(P.nC, P.nP, P.cam_idx, P.pt_idx, P.obs).lm_solve(P.x(), **params)
out = dict(shape=shape, seed=seed, params=params, cost=[t["cost"] for t in tr],
           accepted=[t["accepted"] for t in tr], pcg_iters=[t["pcg_iters"] for t in tr])
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_lm_small.json"), "w") as fh:
    json.dump(out, fh, indent=1)
print(out)

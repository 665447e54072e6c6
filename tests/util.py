"""Shared helpers for the parity tests (problems, tolerances, error norms)."""
import functools

import numpy as np

from megba_b200.synthetic import BAProblem, make_synthetic

# Parity tolerances (BASELINE.md section 5 -- ours; upstream pins none)
TOL = {
    np.dtype(np.float64): dict(rj=1e-12, H=1e-10, dx=1e-8, cost=1e-8),
    np.dtype(np.float32): dict(rj=1e-5, H=1e-4, dx=1e-3, cost=1e-3),
}


def rel_fro(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.linalg.norm(b.ravel())
    return np.linalg.norm((a - b).ravel()) / (den if den > 0 else 1.0)


def rel_rowmax(a, b):
    """max |a-b| relative to the max magnitude of each last-axis row of b."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.abs(b).max(axis=-1, keepdims=True)
    den[den == 0] = 1.0
    return float((np.abs(a - b) / den).max())


@functools.lru_cache(maxsize=None)
def small_problem(seed=1) -> BAProblem:
    return make_synthetic(12, 400, 1700, seed=seed)


@functools.lru_cache(maxsize=None)
def medium_problem(seed=2) -> BAProblem:
    return make_synthetic(49, 7776, 31843, seed=seed)


@functools.lru_cache(maxsize=None)
def bigpoint_problem(seed=3) -> BAProblem:
    """nC = 600 with two points seen by > 256 cameras (multi-slab tiles) and one camera with theta ~ 0."""
    P = make_synthetic(600, 900, 5000, seed=seed)
    rng = np.random.default_rng(seed)
    extra_c, extra_p = [], []
    for p, n in ((7, 600), (431, 300)):
        have = set(P.cam_idx[P.pt_idx == p].tolist())
        cams = [c for c in rng.permutation(600).tolist() if c not in have][: n - len(have)]
        extra_c += cams
        extra_p += [p] * len(cams)
    cam_idx = np.concatenate([P.cam_idx, np.array(extra_c, np.int32)])
    pt_idx = np.concatenate([P.pt_idx, np.array(extra_p, np.int32)])
    # shuffle: the C-ABI accepts observations in any order
    perm = rng.permutation(len(cam_idx))
    cam_idx, pt_idx = cam_idx[perm], pt_idx[perm]
    from megba_b200.synthetic import project
    obs = project(P.cam[cam_idx], P.pt[pt_idx]) + rng.normal(0, 0.5, size=(len(cam_idx), 2))
    return BAProblem(600, 900, len(cam_idx), cam_idx, pt_idx, np.ascontiguousarray(obs), P.cam, P.pt, "bigpoint")


@functools.lru_cache(maxsize=None)
def bal_like_problem(seed=5) -> BAProblem:
    """BAL-like statistics (megba_b200.synthetic._bal_like_graph): heavy-tailed track lengths up to nC/2 (points with a
    warp of their own in the ELL layout, multi-slab points in the tile layout) and coherent camera windows."""
    return make_synthetic(120, 4000, 24000, seed=seed, graph="bal")

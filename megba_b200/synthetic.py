"""Synthetic BAL-shaped bundle-adjustment problems and BAL text I/O.

No BAL file exists in the build container or on the GPU box (no network), so every
configuration in BASELINE.json is reproduced as a synthetic problem with exactly the
(nC, nP, nE) of the real dataset.  Generator spec (adjusted so
that every observation has positive depth):

* seed -> numpy PCG64;  points uniform in a ball of radius 10 around the origin;
* cameras on a sphere r ~ U(25, 35) looking at the origin with a random roll
  (BAL convention: the camera looks down -z, so P_z < 0 for visible points),
  omega = rotation vector of the world->camera rotation, t = -R * centre;
* f ~ U(400, 1200), k1 ~ N(0, (3e-2)^2), k2 ~ N(0, (6e-3)^2) (larger than BAL's
  Ladybug values so the distortion columns of J are exercised);
* degrees: every point gets 2 observations, the remaining nE - 2 nP are dealt
  multinomially; cameras per point drawn without replacement with probability
  proportional to a LogNormal(0, 0.5) popularity weight;
* obs = projection + N(0, 0.5^2) px; initial estimate = truth + noise;
* edges emitted point-major then camera-ascending (the BAL file order).

The file format read/written here is the one parsed by the reference's examples
(examples/BAL_Double_analytical.cpp:71-136): ``nC nP nE`` / ``cam pt u v`` x nE /
9 x nC camera scalars / 3 x nP point scalars.
"""
from __future__ import annotations

import dataclasses

import numpy as np

# (nC, nP, nE) of the BAL problems named in BASELINE.json
BAL_SHAPES = {
    "ladybug-49": (49, 7776, 31843),
    "trafalgar-257": (257, 65132, 225911),
    "venice-1778": (1778, 993923, 5001946),
    "final-13682": (13682, 4456117, 28987644),
}


@dataclasses.dataclass
class BAProblem:
    """Canonical host-side problem: the arrays the C-ABI's ba_create consumes."""

    nC: int
    nP: int
    nE: int
    cam_idx: np.ndarray  # int32 [nE]
    pt_idx: np.ndarray  # int32 [nE]
    obs: np.ndarray  # float64 [nE, 2]
    cam: np.ndarray  # float64 [nC, 9]  initial estimate
    pt: np.ndarray  # float64 [nP, 3]  initial estimate
    label: str = "synthetic"

    def x(self, dtype=np.float64) -> np.ndarray:
        return np.concatenate([self.cam.ravel(), self.pt.ravel()]).astype(dtype)


def project(cam: np.ndarray, X: np.ndarray) -> np.ndarray:
    """BAL projection of points X[n,3] by cameras cam[n,9] (numpy, float64).

    Same model as src/geo/analytical_derivatives.cu:186-195; used only to create
    observations for synthetic data (an independent implementation via Rodrigues'
    matrix form, which doubles as a cross-check of the oracle in tests).
    """
    from scipy.spatial.transform import Rotation

    R = Rotation.from_rotvec(cam[:, :3]).as_matrix()
    P = np.einsum("nij,nj->ni", R, X) + cam[:, 3:6]
    B = -P[:, :2] / P[:, 2:3]
    l2 = (B * B).sum(1)
    rp = 1.0 + cam[:, 7] * l2 + cam[:, 8] * l2 * l2
    return (cam[:, 6] * rp)[:, None] * B


def _bal_like_graph(nC: int, nP: int, nE: int, rng) -> tuple[np.ndarray, np.ndarray]:
    """Observation graph with the statistics of real BAL problems instead of a uniform random one: heavy-tailed track
    lengths (most points seen 2-4 times, a few by hundreds of cameras) and spatially coherent camera sets (cameras are
    ordered along a trajectory; a point is seen by cameras of one window of it, so neighbouring cameras share points
    and camera degrees are skewed).  Returns (pt_idx, cam_idx), point-major, camera-ascending inside a point."""
    raw = rng.pareto(1.6, nP) + 0.05
    deg = 2 + rng.multinomial(nE - 2 * nP, raw / raw.sum())
    cap = max(2, nC // 2)
    over = deg > cap
    while over.any():
        excess = int((deg[over] - cap).sum())
        deg[over] = cap
        room = np.flatnonzero(deg < cap)
        w = raw[room] / raw[room].sum()
        deg[room] += rng.multinomial(excess, w)
        over = deg > cap
    pt_idx = np.repeat(np.arange(nP, dtype=np.int64), deg)
    start = np.concatenate([[0], np.cumsum(deg)[:-1]])
    # cameras of a point: c0 + cumulative gaps (>= 1) modulo nC; mean gap <= 3 and span < nC keep them distinct and close
    mean_gap = np.minimum(3.0, (nC - 1) / np.maximum(deg, 1) * 0.9)
    gaps = 1 + np.floor(rng.exponential(np.repeat(np.maximum(mean_gap - 1.0, 0.0), deg))).astype(np.int64)
    gaps[start] = 0
    off = np.cumsum(gaps)
    off -= np.repeat(off[start], deg)
    span = off[start + deg - 1]
    bad = np.repeat(span >= nC, deg)  # fall back to consecutive cameras for the few windows that came out too wide
    if bad.any():
        pos = np.arange(len(off)) - np.repeat(start, deg)
        off[bad] = pos[bad]
    hot = rng.random(nP) < 0.3  # a third of the points sit in a few dense regions of the trajectory (skewed camera degrees)
    c0 = np.where(hot, (rng.integers(0, 4, nP) * (nC // 4) + rng.integers(0, max(nC // 16, 1), nP)) % nC, rng.integers(0, nC, nP))
    cam_idx = (np.repeat(c0, deg) + off) % nC
    key = pt_idx * nC + cam_idx
    order = np.argsort(key, kind="stable")
    return pt_idx[order].astype(np.int32), cam_idx[order].astype(np.int32)


def make_synthetic(nC: int, nP: int, nE: int, seed: int = 42, noise_px: float = 0.5,
                   label: str | None = None, graph: str = "uniform") -> BAProblem:
    from scipy.spatial.transform import Rotation

    if nE < 2 * nP:
        raise ValueError("need at least two observations per point")
    rng = np.random.Generator(np.random.PCG64(seed))
    # --- ground truth geometry
    d = rng.normal(size=(nP, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    X = d * (10.0 * rng.random(nP) ** (1.0 / 3.0))[:, None]
    cdir = rng.normal(size=(nC, 3))
    cdir /= np.linalg.norm(cdir, axis=1, keepdims=True)
    centre = cdir * rng.uniform(25.0, 35.0, size=(nC, 1))
    zc = cdir  # camera +z points away from the scene: looks down -z at the origin
    helper = np.where(np.abs(zc[:, 2:3]) < 0.9, np.array([[0.0, 0.0, 1.0]]), np.array([[1.0, 0.0, 0.0]]))
    xc = np.cross(helper, zc)
    xc /= np.linalg.norm(xc, axis=1, keepdims=True)
    yc = np.cross(zc, xc)
    roll = rng.uniform(0, 2 * np.pi, size=nC)
    xr = np.cos(roll)[:, None] * xc + np.sin(roll)[:, None] * yc
    yr = -np.sin(roll)[:, None] * xc + np.cos(roll)[:, None] * yc
    Rw2c = np.stack([xr, yr, zc], axis=1)  # rows = camera axes in world coordinates
    omega = Rotation.from_matrix(Rw2c).as_rotvec()
    t = -np.einsum("nij,nj->ni", Rw2c, centre)
    f = rng.uniform(400.0, 1200.0, size=nC)
    k1 = rng.normal(0.0, 3e-2, size=nC)
    k2 = rng.normal(0.0, 6e-3, size=nC)
    cam_true = np.concatenate([omega, t, f[:, None], k1[:, None], k2[:, None]], axis=1)
    # --- graph
    if graph == "bal":
        pt_idx, cam_idx = _bal_like_graph(nC, nP, nE, rng)
        return _finish(nC, nP, nE, cam_idx, pt_idx, cam_true, X, rng, noise_px,
                       label or f"synthetic-bal-{nC}-{nP}-{nE}-seed{seed}")
    deg = np.full(nP, 2, dtype=np.int64)
    extra = nE - 2 * nP
    if extra:
        deg += rng.multinomial(extra, np.full(nP, 1.0 / nP))
    over = deg > nC
    while over.any():  # tiny camera counts: move the excess to points with room
        excess = int((deg[over] - nC).sum())
        deg[over] = nC
        room = np.flatnonzero(deg < nC)
        add = rng.multinomial(excess, np.full(room.size, 1.0 / room.size))
        deg[room] += add
        over = deg > nC
    pt_idx = np.repeat(np.arange(nP, dtype=np.int64), deg)
    wts = rng.lognormal(0.0, 0.5, size=nC)
    cum = np.cumsum(wts / wts.sum())
    cam_idx = np.minimum(np.searchsorted(cum, rng.random(nE)), nC - 1).astype(np.int64)
    for _ in range(1000):
        key = pt_idx * nC + cam_idx
        order = np.argsort(key, kind="stable")
        ks = key[order]
        dup = np.flatnonzero(ks[1:] == ks[:-1]) + 1
        if dup.size == 0:
            break
        pos = order[dup]
        cam_idx[pos] = np.minimum(np.searchsorted(cum, rng.random(pos.size)), nC - 1)
    else:  # pragma: no cover
        raise RuntimeError("could not de-duplicate (point, camera) pairs")
    key = pt_idx * nC + cam_idx
    order = np.argsort(key, kind="stable")
    pt_idx = pt_idx[order].astype(np.int32)
    cam_idx = cam_idx[order].astype(np.int32)
    return _finish(nC, nP, nE, cam_idx, pt_idx, cam_true, X, rng, noise_px,
                   label or f"synthetic-{nC}-{nP}-{nE}-seed{seed}")


def _finish(nC, nP, nE, cam_idx, pt_idx, cam_true, X, rng, noise_px, label) -> BAProblem:
    # --- observations and the perturbed initial estimate
    obs = project(cam_true[cam_idx], X[pt_idx]) + rng.normal(0.0, noise_px, size=(nE, 2))
    cam0 = cam_true.copy()
    cam0[:, :3] += rng.normal(0.0, 1e-3, size=(nC, 3))
    cam0[:, 3:6] += rng.normal(0.0, 5e-2, size=(nC, 3))
    cam0[:, 6] *= 1.0 + rng.normal(0.0, 1e-2, size=nC)
    pt0 = X + rng.normal(0.0, 0.05, size=(nP, 3))
    return BAProblem(nC, nP, nE, cam_idx, pt_idx, np.ascontiguousarray(obs), cam0, pt0, label)


def make_named(name: str, seed: int = 42) -> BAProblem:
    """`venice-1778` = uniform random graph of that shape (the headline workload since round 1); `venice-1778:bal` = the
    same shape with BAL-like statistics (heavy-tailed tracks, coherent camera windows)."""
    base, _, kind = name.partition(":")
    base, _, div = base.partition("/")  # "venice-1778/8": one eighth of the points and observations, all cameras (a shard's size)
    nC, nP, nE = BAL_SHAPES[base]
    if div:
        nP, nE = nP // int(div), nE // int(div)
        base = f"{base}/{div}"
    if kind == "bal":
        return make_synthetic(nC, nP, nE, seed=seed, label=f"synthetic-bal:{base}", graph="bal")
    return make_synthetic(nC, nP, nE, seed=seed, label=f"synthetic:{base}")


def write_bal(path: str, prob: BAProblem) -> None:
    """BAL text format (examples/BAL_Double_analytical.cpp:71-136 reads this)."""
    with open(path, "w") as fh:
        fh.write(f"{prob.nC} {prob.nP} {prob.nE}\n")
        for c, p, (u, v) in zip(prob.cam_idx, prob.pt_idx, prob.obs):
            fh.write(f"{int(c)} {int(p)} {float(u)!r} {float(v)!r}\n")
        for val in prob.cam.ravel():
            fh.write(f"{float(val)!r}\n")
        for val in prob.pt.ravel():
            fh.write(f"{float(val)!r}\n")


def read_bal(path: str) -> BAProblem:
    with open(path) as fh:
        tok = fh.read().split()
    nC, nP, nE = int(tok[0]), int(tok[1]), int(tok[2])
    body = np.array(tok[3:3 + 4 * nE], dtype=np.float64).reshape(nE, 4)
    rest = np.array(tok[3 + 4 * nE:3 + 4 * nE + 9 * nC + 3 * nP], dtype=np.float64)
    return BAProblem(nC, nP, nE, body[:, 0].astype(np.int32), body[:, 1].astype(np.int32),
                     np.ascontiguousarray(body[:, 2:4]), rest[:9 * nC].reshape(nC, 9).copy(),
                     rest[9 * nC:].reshape(nP, 3).copy(), label=f"bal:{path}")

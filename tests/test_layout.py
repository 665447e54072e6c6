"""Host logic of ba_create without a GPU: the observation layout (point-major tiles, camera-major chunks, point
partition) returned by the host-only C-ABI entry ba_layout.  These are the invariants the kernels rely on."""
import numpy as np
import pytest

from megba_b200 import capi
from megba_b200.synthetic import make_synthetic
from util import bal_like_problem, bigpoint_problem, small_problem

TE, CM_CHUNK = 256, 1024


def _problems():
    rng = np.random.default_rng(5)
    ragged = make_synthetic(7, 300, 900, seed=9)
    keep = rng.random(ragged.nE) < 0.6  # leaves points with 0 / 1 observations
    return {
        "small": (small_problem().nC, small_problem().nP, small_problem().cam_idx, small_problem().pt_idx),
        "bigpoint": (bigpoint_problem().nC, bigpoint_problem().nP, bigpoint_problem().cam_idx, bigpoint_problem().pt_idx),
        "ragged": (ragged.nC, ragged.nP, ragged.cam_idx[keep], ragged.pt_idx[keep]),
        "single_obs_points": (5, 1000, rng.integers(0, 5, 1000).astype(np.int32), np.arange(1000, dtype=np.int32)),
        "empty": (3, 4, np.zeros(0, np.int32), np.zeros(0, np.int32)),
        "bal": (bal_like_problem().nC, bal_like_problem().nP, bal_like_problem().cam_idx, bal_like_problem().pt_idx),
    }


@pytest.mark.parametrize("name", ["small", "bigpoint", "ragged", "single_obs_points", "empty", "bal"])
@pytest.mark.parametrize("ws", [1, 2, 3])
def test_layout_invariants(name, ws):
    nC, nP, cam_idx, pt_idx = _problems()[name]
    nE = len(cam_idx)
    deg = np.bincount(pt_idx, minlength=nP)
    cuts = capi.partition_points(deg, ws)
    seen = np.zeros(nE, np.int32)
    for rank in range(ws):
        L = capi.layout(nC, nP, cam_idx, pt_idx, rank, ws)
        tile, eos, pt_slot = L["tile"], L["edge_of_slot"], L["pt_slot"]
        assert L["pt_begin"] == cuts[rank] and len(pt_slot) == cuts[rank + 1] - cuts[rank]
        # tile index == slab index; the kernels address a CTA's slots from blockIdx alone
        assert np.array_equal(tile[:, 0], TE * np.arange(len(tile)))
        assert len(eos) == TE * len(tile)
        local = eos[eos >= 0]
        seen[local] += 1
        assert np.all((pt_idx[local] >= cuts[rank]) & (pt_idx[local] < cuts[rank + 1]))
        assert len(local) == deg[cuts[rank]:cuts[rank + 1]].sum() == len(L["cm_slot"])
        t = 0
        while t < len(tile):
            s0, ne, p0, npts = tile[t]
            if ne > TE:  # one point spanning several slabs, followed by its empty tiles
                assert npts == 1
                extra = (ne + TE - 1) // TE - 1
                assert np.all(tile[t + 1:t + 1 + extra, 1] == 0) and np.all(tile[t + 1:t + 1 + extra, 3] == 0)
                span = ne
            else:
                assert 0 < npts <= TE or (ne == 0 and npts == 0)
                extra, span = 0, ne
            slots = np.arange(s0, s0 + span)
            assert np.all(eos[slots] >= 0) and np.all(eos[s0 + span:s0 + TE * (extra + 1)] == -1)
            # slots of a tile: points in order, each point contiguous from pt_slot, camera-ascending inside a point
            lp = L["slot_lp"][slots]
            assert np.array_equal(pt_idx[eos[slots]] - cuts[rank] - p0, lp) and np.all(np.diff(lp) >= 0)
            assert lp.size == 0 or lp.max() < max(npts, 1)
            assert np.array_equal(L["slot_cam"][slots], cam_idx[eos[slots]])
            for q in range(npts):
                a = pt_slot[p0 + q]
                d = deg[cuts[rank] + p0 + q]
                assert s0 <= a and a + d <= s0 + max(span, 0) + (0 if d else TE)
                if d:
                    assert np.all(pt_idx[eos[a:a + d]] == cuts[rank] + p0 + q)
                    assert np.all(np.diff(cam_idx[eos[a:a + d]]) >= 0)
            t += 1 + extra
        # every local point belongs to exactly one tile
        assert tile[:, 3].sum() == len(pt_slot)
        # camera-major view: chunks partition [0, nEl), one camera each, at most CM_CHUNK entries
        chunk, cm = L["chunk"], L["cm_slot"]
        pos = 0
        for c, b, e, _ in chunk:
            assert b == pos and b < e <= b + CM_CHUNK
            assert np.all(L["slot_cam"][cm[b:e]] == c)
            pos = e
        assert pos == len(cm)
        assert np.array_equal(np.sort(cm), np.flatnonzero(eos >= 0))
        assert np.all(np.diff(chunk[:, 0]) >= 0) if len(chunk) else True
    assert np.all(seen == 1)  # every observation lives on exactly one rank, in exactly one slot


def test_layout_rejects_bad_input():
    with pytest.raises(RuntimeError):
        capi.layout(2, 2, np.array([0, 5], np.int32), np.array([0, 1], np.int32))  # camera index out of range
    with pytest.raises(RuntimeError):
        capi.layout(1 << 24, 2, np.array([0], np.int32), np.array([0], np.int32))  # camera count beyond the packed index


@pytest.mark.parametrize("name", ["small", "bigpoint", "ragged", "single_obs_points", "empty", "bal"])
@pytest.mark.parametrize("ws", [1, 2, 3])
def test_ell_layout_invariants(name, ws):
    """The default (ELL) layout: device order sorted by observation count, heavy points in warps of their own,
    slices of 32 points with one row per observation rank, every observation in exactly one slot of one rank."""
    nC, nP, cam_idx, pt_idx = _problems()[name]
    nE = len(cam_idx)
    deg = np.bincount(pt_idx, minlength=nP)
    cuts = capi.partition_points(deg, ws)
    seen = np.zeros(nE, np.int32)
    for rank in range(ws):
        L = capi.layout_ell(nC, nP, cam_idx, pt_idx, rank, ws)
        warp, perm, eos, scam = L["warp"], L["perm"], L["edge_of_slot"], L["slot_cam"]
        nPl = cuts[rank + 1] - cuts[rank]
        assert L["pt_begin"] == cuts[rank] and len(perm) == nPl
        assert sorted(perm.tolist()) == list(range(nPl))  # a permutation of the local points
        ldeg = deg[cuts[rank] + perm] if nPl else np.zeros(0, np.int64)
        assert np.all(np.diff(ldeg) <= 0)  # descending observation count
        assert L["n_heavy"] == int((ldeg > 32).sum())
        assert len(eos) == 32 * L["n_rows"] and len(scam) == len(eos)
        row = 0
        next_pt = 0
        for w, (row0, D, pt0, kw) in enumerate(warp):
            kind, npts = kw & 0xFF, (kw >> 8) & 0xFF
            assert row0 == row and pt0 == next_pt  # rows and points are consumed in order, without gaps
            if kind == 1:  # heavy point: lane = observation
                assert w < L["n_heavy"] and npts == 1 and D == -(-ldeg[pt0] // 32)
                slots = eos[32 * row0: 32 * (row0 + D)]
                assert np.all(slots[: ldeg[pt0]] >= 0) and np.all(slots[ldeg[pt0]:] == -1)
                assert np.all(pt_idx[slots[: ldeg[pt0]]] == cuts[rank] + perm[pt0])
            else:  # slice: lane = point, row r = r-th observation of each point
                assert 1 <= npts <= 32 and D == ldeg[pt0]
                blk = eos[32 * row0: 32 * (row0 + D)].reshape(D, 32)
                for lane in range(32):
                    col = blk[:, lane]
                    d = ldeg[pt0 + lane] if lane < npts else 0
                    assert np.all(col[:d] >= 0) and np.all(col[d:] == -1)
                    if d:
                        assert np.all(pt_idx[col[:d]] == cuts[rank] + perm[pt0 + lane])
                        assert np.all(np.diff(cam_idx[col[:d]]) >= 0)  # camera-ascending inside a point
            row += D
            next_pt += npts
        assert row == L["n_rows"] and next_pt == nPl
        valid = eos >= 0
        assert np.array_equal(scam[valid], cam_idx[eos[valid]]) and np.all(scam[~valid] == -1)
        np.add.at(seen, eos[valid], 1)
        # camera-major view: chunks partition it, every entry points at a slot of that camera
        chunk, cms = L["chunk"], L["cm_slot"]
        assert len(cms) == valid.sum()
        pos = 0
        for cam, beg, end, _ in chunk:
            assert beg == pos and 0 < end - beg <= CM_CHUNK
            assert np.all(scam[cms[beg:end]] == cam)
            pos = end
        assert pos == len(cms)
        assert len(np.unique(cms)) == len(cms)
    assert np.all(seen == 1)

"""ctypes binding of the C-ABI in include/megba_b200.h (test / benchmark harness).

This module only forwards to libmegba_b200.so.  There is deliberately no Python or
CPU implementation behind it: if the library is missing or no CUDA device is present
the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmegba_b200.so")

BA_F64, BA_F32 = 0, 1
BA_EXPLICIT, BA_IMPLICIT = 0, 1
LAYOUT_AUTO, LAYOUT_TILE, LAYOUT_ELL = 0, 1, 2
STORE_AUTO, STORE_W, STORE_J, STORE_FACT = 0, 1, 2, 3
BUF_HPP, BUF_HLL, BUF_W, BUF_J, BUF_G, BUF_DX = range(6)

# every symbol include/megba_b200.h declares (tests check the export list against the header)
SYMBOLS = [
    "ba_abi_version", "ba_partition_points", "ba_create", "ba_destroy", "ba_last_error",
    "ba_comm_unique_id", "ba_comm_init", "ba_comm_init_local", "ba_set_params", "ba_get_params", "ba_set_information", "ba_set_fixed",
    "ba_set_residual_program", "ba_residual_program_mode", "ba_residual_program_source",
    "ba_linearize", "ba_process_diag", "ba_solve", "ba_solve_async", "ba_solve_wait",
    "ba_set_params_device", "ba_get_params_device", "ba_step_norms", "ba_eval_trial",
    "ba_accept_step", "ba_reject_step", "ba_grad_inf_norm", "ba_lm_solve", "ba_read", "ba_write_dx",
    "ba_eval_edges", "ba_schur_apply", "ba_set_timing", "ba_reset_stats", "ba_get_stats",
    "ba_algorithmic_bytes", "ba_stream", "ba_set_persistent", "ba_layout", "ba_layout_ell",
]


class Options(C.Structure):
    _fields_ = [("dtype", C.c_int32), ("compute_kind", C.c_int32), ("device", C.c_int32),
                ("rank", C.c_int32), ("world_size", C.c_int32), ("layout", C.c_int32),
                ("operator_store", C.c_int32), ("reserved", C.c_int32 * 1)]


class PcgOptions(C.Structure):
    _fields_ = [("max_iter", C.c_int32), ("reserved", C.c_int32), ("tol", C.c_double),
                ("refuse_ratio", C.c_double)]


class LmOptions(C.Structure):
    _fields_ = [("max_iter", C.c_int32), ("reserved", C.c_int32), ("initial_region", C.c_double),
                ("epsilon1", C.c_double), ("epsilon2", C.c_double)]


class LmTrace(C.Structure):
    _fields_ = [("cost", C.c_double), ("accepted", C.c_int32), ("pcg_iters", C.c_int32),
                ("region", C.c_double), ("dx_norm", C.c_double), ("rho", C.c_double),
                ("elapsed_ms", C.c_double)]


class Stats(C.Structure):
    _fields_ = [("launches", C.c_int64), ("lm_iters", C.c_int64), ("pcg_iters", C.c_int64),
                ("schur_apply_calls", C.c_int64), ("schur_apply_ms", C.c_double),
                ("linearize_calls", C.c_int64), ("linearize_ms", C.c_double),
                ("pcg_loop_ms", C.c_double), ("solve_ms", C.c_double), ("pcg_cam_ms", C.c_double),
                ("schur_apply_timed", C.c_double), ("pcg_cam_timed", C.c_double), ("reserved", C.c_double * 1)]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}


_lib = None


def load():
    """dlopen the product library; raises if it was not built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(megba_b200 has no CPU or pure-Python fallback)")
        L = C.CDLL(LIB_PATH)
        L.ba_last_error.restype = C.c_char_p
        L.ba_last_error.argtypes = [C.c_void_p]
        L.ba_stream.restype = C.c_void_p
        L.ba_stream.argtypes = [C.c_void_p]
        L.ba_destroy.restype = None
        L.ba_destroy.argtypes = [C.c_void_p]
        L.ba_create.argtypes = [C.POINTER(Options), C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.POINTER(C.c_void_p)]
        L.ba_partition_points.argtypes = [C.c_int64, C.c_void_p, C.c_int32, C.c_void_p]
        for name in ("ba_comm_init", "ba_set_params", "ba_get_params", "ba_set_information", "ba_linearize",
                     "ba_solve", "ba_solve_async", "ba_solve_wait", "ba_set_params_device", "ba_get_params_device",
                     "ba_step_norms", "ba_eval_trial", "ba_accept_step", "ba_reject_step",
                     "ba_grad_inf_norm", "ba_lm_solve", "ba_read", "ba_write_dx", "ba_eval_edges",
                     "ba_schur_apply", "ba_set_timing", "ba_reset_stats", "ba_get_stats",
                     "ba_algorithmic_bytes"):
            getattr(L, name).restype = C.c_int
        L.ba_process_diag.argtypes = [C.c_void_p, C.c_double, C.c_int]
        L.ba_process_diag.restype = C.c_int
        _lib = L
    return _lib


def residual_program_source(code, consts, n_slots, out, dtype=np.float64, compile=True):
    """Host-only: (CUDA source generated for the program, size of the sm_100a image NVRTC built from it or None)."""
    L = load()
    code = np.ascontiguousarray(code, np.uint8).reshape(-1, 6)
    k = np.ascontiguousarray(consts, np.float64)
    o = np.ascontiguousarray(out, np.int32)
    n = C.c_int64()
    cub = C.c_int64()
    args = (int(code.shape[0]), _p(code), int(k.size), _p(k) if k.size else None, int(n_slots), _p(o),
            0 if np.dtype(dtype) == np.float64 else 1)
    rc = L.ba_residual_program_source(*args, None, C.c_int64(0), C.byref(n), None)
    if rc != 0:
        raise RuntimeError(f"ba_residual_program_source failed ({rc}): {L.ba_last_error(None).decode()}")
    buf = C.create_string_buffer(n.value + 1)
    rc = L.ba_residual_program_source(*args, buf, C.c_int64(n.value + 1), C.byref(n), C.byref(cub) if compile else None)
    if rc != 0:
        raise RuntimeError(f"ba_residual_program_source failed ({rc}): {L.ba_last_error(None).decode()}")
    return buf.value.decode(), (cub.value if compile else None)


def partition_points(pt_degree: np.ndarray, world_size: int) -> np.ndarray:
    deg = np.ascontiguousarray(pt_degree, dtype=np.int64)
    out = np.zeros(world_size + 1, dtype=np.int64)
    rc = load().ba_partition_points(len(deg), deg.ctypes.data_as(C.c_void_p), world_size,
                                    out.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise RuntimeError(f"ba_partition_points failed: {rc}")
    return out


class LayoutSizes(C.Structure):
    _fields_ = [("n_tiles", C.c_int64), ("n_slots", C.c_int64), ("n_local_edges", C.c_int64),
                ("n_local_points", C.c_int64), ("n_chunks", C.c_int64), ("pt_begin", C.c_int64)]


def layout(nC, nP, cam_idx, pt_idx, rank=0, world_size=1):
    """Host-only (no GPU): the observation layout ba_create builds for (rank, world_size), as a dict of arrays."""
    L = load()
    cam_idx = np.ascontiguousarray(cam_idx, np.int32)
    pt_idx = np.ascontiguousarray(pt_idx, np.int32)
    nE = len(cam_idx)
    opt = Options(0, 0, 0, rank, world_size)
    sz = LayoutSizes()
    null = C.c_void_p(None)

    def call(*arrays):
        rc = L.ba_layout(C.byref(opt), C.c_int64(nC), C.c_int64(nP), C.c_int64(nE), C.c_void_p(cam_idx.ctypes.data),
                         C.c_void_p(pt_idx.ctypes.data), C.byref(sz), *arrays)
        if rc != 0:
            raise RuntimeError(f"ba_layout failed: {rc}")

    call(*([null] * 7))
    out = dict(tile=np.zeros((sz.n_tiles, 4), np.int32), edge_of_slot=np.zeros(sz.n_slots, np.int64),
               slot_cam=np.zeros(sz.n_slots, np.int32), slot_lp=np.zeros(sz.n_slots, np.int32),
               pt_slot=np.zeros(sz.n_local_points, np.int32), chunk=np.zeros((sz.n_chunks, 4), np.int32),
               cm_slot=np.zeros(sz.n_local_edges, np.int32))
    call(*(C.c_void_p(out[k].ctypes.data) for k in ("tile", "edge_of_slot", "slot_cam", "slot_lp", "pt_slot", "chunk", "cm_slot")))
    out["pt_begin"] = int(sz.pt_begin)
    return out


class EllSizes(C.Structure):
    _fields_ = [("n_warps", C.c_int64), ("n_rows", C.c_int64), ("n_slots", C.c_int64), ("n_local_edges", C.c_int64),
                ("n_local_points", C.c_int64), ("n_chunks", C.c_int64), ("pt_begin", C.c_int64), ("n_heavy", C.c_int64)]


def layout_ell(nC, nP, cam_idx, pt_idx, rank=0, world_size=1):
    """Host-only (no GPU): the ELL observation layout (the default of ba_create), as a dict of arrays."""
    L = load()
    cam_idx = np.ascontiguousarray(cam_idx, np.int32)
    pt_idx = np.ascontiguousarray(pt_idx, np.int32)
    nE = len(cam_idx)
    opt = Options(0, 0, 0, rank, world_size)
    sz = EllSizes()
    null = C.c_void_p(None)

    def call(*arrays):
        rc = L.ba_layout_ell(C.byref(opt), C.c_int64(nC), C.c_int64(nP), C.c_int64(nE), C.c_void_p(cam_idx.ctypes.data),
                             C.c_void_p(pt_idx.ctypes.data), C.byref(sz), *arrays)
        if rc != 0:
            raise RuntimeError(f"ba_layout_ell failed: {rc}")

    call(*([null] * 6))
    out = dict(warp=np.zeros((sz.n_warps, 4), np.int32), perm=np.zeros(sz.n_local_points, np.int32),
               edge_of_slot=np.zeros(sz.n_slots, np.int64), slot_cam=np.zeros(sz.n_slots, np.int32),
               chunk=np.zeros((sz.n_chunks, 4), np.int32), cm_slot=np.zeros(sz.n_local_edges, np.int32))
    call(*(C.c_void_p(out[k].ctypes.data) for k in ("warp", "perm", "edge_of_slot", "slot_cam", "chunk", "cm_slot")))
    out["pt_begin"] = int(sz.pt_begin)
    out["n_heavy"] = int(sz.n_heavy)
    out["n_rows"] = int(sz.n_rows)
    return out


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


class Solver:
    """Thin owner of one ba_solver handle; method names follow the C-ABI."""

    def __init__(self, nC, nP, cam_idx, pt_idx, obs, dtype=np.float64, implicit=False, device=0, rank=0,
                 world_size=1, layout=LAYOUT_AUTO, store=STORE_AUTO):
        self.L = load()
        self.dtype = np.dtype(dtype)
        self.nC, self.nP, self.nE = int(nC), int(nP), int(len(cam_idx))
        self.implicit = bool(implicit)
        opt = Options(BA_F64 if self.dtype == np.float64 else BA_F32, BA_IMPLICIT if implicit else BA_EXPLICIT,
                      device, rank, world_size, layout, store)
        cam_idx = np.ascontiguousarray(cam_idx, dtype=np.int32)
        pt_idx = np.ascontiguousarray(pt_idx, dtype=np.int32)
        obs = np.ascontiguousarray(obs, dtype=self.dtype)
        self.h = C.c_void_p()
        rc = self.L.ba_create(C.byref(opt), self.nC, self.nP, self.nE, _p(cam_idx), _p(pt_idx), _p(obs),
                              C.byref(self.h))
        if rc != 0:
            raise RuntimeError(f"ba_create failed ({rc}): {self.L.ba_last_error(None).decode()}")

    def close(self):
        if getattr(self, "h", None) is not None and self.h.value:
            self.L.ba_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"{what} failed ({rc}): {self.L.ba_last_error(self.h).decode()}")

    def _a(self, a):
        return np.ascontiguousarray(a, dtype=self.dtype)

    # -- communicator
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        rc = load().ba_comm_unique_id(buf)
        if rc != 0:
            raise RuntimeError(f"ba_comm_unique_id failed: {rc}")
        return buf.raw

    @staticmethod
    def comm_init_local(solvers):
        """In-process communicator over `solvers` (rank r = solvers[r]); afterwards drive each solver from its own thread."""
        arr = (C.c_void_p * len(solvers))(*[s.h.value for s in solvers])
        rc = load().ba_comm_init_local(arr, len(solvers))
        if rc != 0:
            raise RuntimeError(f"ba_comm_init_local failed: {rc}")

    def comm_init(self, uid: bytes):
        self._chk(self.L.ba_comm_init(self.h, C.c_char_p(uid)), "ba_comm_init")

    # -- parameters
    def set_params(self, cam, pt):
        cam, pt = self._a(cam), self._a(pt)
        self._chk(self.L.ba_set_params(self.h, _p(cam), _p(pt)), "ba_set_params")

    def set_params_ptr(self, cam_ptr: int, pt_ptr: int):
        self._chk(self.L.ba_set_params(self.h, C.c_void_p(cam_ptr), C.c_void_p(pt_ptr)), "ba_set_params")

    def get_params_ptr(self, cam_ptr: int, pt_ptr: int):
        self._chk(self.L.ba_get_params(self.h, C.c_void_p(cam_ptr), C.c_void_p(pt_ptr)), "ba_get_params")

    def get_params(self):
        cam = np.empty((self.nC, 9), self.dtype)
        pt = np.empty((self.nP, 3), self.dtype)
        self._chk(self.L.ba_get_params(self.h, _p(cam), _p(pt)), "ba_get_params")
        return cam, pt

    def set_information(self, info):
        info = None if info is None else self._a(info)
        self._chk(self.L.ba_set_information(self.h, _p(info)), "ba_set_information")

    def set_fixed(self, cam_fixed=None, pt_fixed=None):
        cf = None if cam_fixed is None else np.ascontiguousarray(cam_fixed, np.uint8)
        pf = None if pt_fixed is None else np.ascontiguousarray(pt_fixed, np.uint8)
        self._chk(self.L.ba_set_fixed(self.h, _p(cf), _p(pf)), "ba_set_fixed")

    def set_residual_program(self, code=None, consts=(), n_slots=0, out=(0, 0)):
        """code: [n_instr, 6] uint8 {op, dst, a, b, c, d} (include/megba_b200.h); None restores the analytical model."""
        if code is None:
            self._chk(self.L.ba_set_residual_program(self.h, 0, None, 0, None, 0, None), "ba_set_residual_program")
            return
        code = np.ascontiguousarray(code, np.uint8).reshape(-1, 6)
        k = np.ascontiguousarray(consts, np.float64)
        o = np.ascontiguousarray(out, np.int32)
        self._chk(self.L.ba_set_residual_program(self.h, int(code.shape[0]), _p(code), int(k.size), _p(k) if k.size else None,
                                                 int(n_slots), _p(o)), "ba_set_residual_program")

    def residual_program_mode(self) -> int:
        """0 analytical kernels, 1 device interpreter, 2 NVRTC-generated kernels."""
        return int(self.L.ba_residual_program_mode(self.h))

    # -- hot path
    def linearize(self) -> float:
        c = C.c_double()
        self._chk(self.L.ba_linearize(self.h, C.byref(c)), "ba_linearize")
        return c.value

    def process_diag(self, region: float, recover: bool = False):
        self._chk(self.L.ba_process_diag(self.h, C.c_double(region), int(recover)), "ba_process_diag")

    def solve(self, max_iter=100, tol=1e-1, refuse_ratio=1.0):
        o = PcgOptions(max_iter, 0, tol, refuse_ratio)
        it, done = C.c_int32(), C.c_int32()
        self._chk(self.L.ba_solve(self.h, C.byref(o), C.byref(it), C.byref(done)), "ba_solve")
        return it.value, bool(done.value)

    def solve_async(self, max_iter=100, tol=1e-1, refuse_ratio=1.0):
        o = PcgOptions(max_iter, 0, tol, refuse_ratio)
        self._chk(self.L.ba_solve_async(self.h, C.byref(o)), "ba_solve_async")

    def solve_wait(self):
        it, done = C.c_int32(), C.c_int32()
        self._chk(self.L.ba_solve_wait(self.h, C.byref(it), C.byref(done)), "ba_solve_wait")
        return it.value, bool(done.value)

    def set_params_device(self, cam_ptr: int, pt_ptr: int, stream: int = 0):
        self._chk(self.L.ba_set_params_device(self.h, C.c_void_p(cam_ptr), C.c_void_p(pt_ptr), C.c_void_p(stream)),
                  "ba_set_params_device")

    def get_params_device(self, cam_ptr: int, pt_ptr: int, stream: int = 0):
        self._chk(self.L.ba_get_params_device(self.h, C.c_void_p(cam_ptr), C.c_void_p(pt_ptr), C.c_void_p(stream)),
                  "ba_get_params_device")

    def step_norms(self):
        a, b = C.c_double(), C.c_double()
        self._chk(self.L.ba_step_norms(self.h, C.byref(a), C.byref(b)), "ba_step_norms")
        return a.value, b.value

    def eval_trial(self):
        a, b = C.c_double(), C.c_double()
        self._chk(self.L.ba_eval_trial(self.h, C.byref(a), C.byref(b)), "ba_eval_trial")
        return a.value, b.value

    def accept_step(self):
        self._chk(self.L.ba_accept_step(self.h), "ba_accept_step")

    def reject_step(self):
        self._chk(self.L.ba_reject_step(self.h), "ba_reject_step")

    def grad_inf_norm(self) -> float:
        g = C.c_double()
        self._chk(self.L.ba_grad_inf_norm(self.h, C.byref(g)), "ba_grad_inf_norm")
        return g.value

    def lm_solve(self, lm_max_iter=20, initial_region=1e3, eps1=1.0, eps2=1e-10, pcg_max_iter=100, pcg_tol=1e-1,
                 pcg_refuse=1.0, verbose=False):
        lm = LmOptions(lm_max_iter, 0, initial_region, eps1, eps2)
        pcg = PcgOptions(pcg_max_iter, 0, pcg_tol, pcg_refuse)
        cap = lm_max_iter + 2
        tr = (LmTrace * cap)()
        nt = C.c_int32()
        self._chk(self.L.ba_lm_solve(self.h, C.byref(lm), C.byref(pcg), tr, cap, C.byref(nt), int(verbose)),
                  "ba_lm_solve")
        return [dict(cost=t.cost, accepted=t.accepted, pcg_iters=t.pcg_iters, region=t.region,
                     dx_norm=t.dx_norm, rho=t.rho, elapsed_ms=t.elapsed_ms) for t in tr[:min(nt.value, cap)]]

    # -- inspection
    def read(self, which):
        shape = {BUF_HPP: (self.nC, 9, 9), BUF_HLL: (self.nP, 3, 3), BUF_W: (self.nE, 9, 3),
                 BUF_J: (self.nE, 2, 12), BUF_G: (9 * self.nC + 3 * self.nP,),
                 BUF_DX: (9 * self.nC + 3 * self.nP,)}[which]
        out = np.empty(shape, self.dtype)
        self._chk(self.L.ba_read(self.h, which, _p(out)), "ba_read")
        return out

    def write_dx(self, dx):
        dx = self._a(dx)
        self._chk(self.L.ba_write_dx(self.h, _p(dx)), "ba_write_dx")

    def eval_edges(self, jac=True):
        r = np.empty((self.nE, 2), self.dtype)
        J = np.empty((self.nE, 2, 12), self.dtype) if jac else None
        self._chk(self.L.ba_eval_edges(self.h, _p(r), _p(J)), "ba_eval_edges")
        return r, J

    def schur_apply(self, p):
        p = self._a(p)
        q = np.empty(9 * self.nC, self.dtype)
        self._chk(self.L.ba_schur_apply(self.h, _p(p), _p(q)), "ba_schur_apply")
        return q

    def set_timing(self, on):
        """0/False off, 1/True every launch, N > 1 every N-th PCG iteration."""
        self._chk(self.L.ba_set_timing(self.h, int(on)), "ba_set_timing")

    def set_persistent(self, on: bool):
        self._chk(self.L.ba_set_persistent(self.h, int(bool(on))), "ba_set_persistent")

    def reset_stats(self):
        self._chk(self.L.ba_reset_stats(self.h), "ba_reset_stats")

    def stats(self) -> dict:
        s = Stats()
        self._chk(self.L.ba_get_stats(self.h, C.byref(s)), "ba_get_stats")
        return s.asdict()

    def algorithmic_bytes(self):
        a, b = C.c_double(), C.c_double()
        self._chk(self.L.ba_algorithmic_bytes(self.h, C.byref(a), C.byref(b)), "ba_algorithmic_bytes")
        return a.value, b.value

    def stream(self) -> int:
        return int(self.L.ba_stream(self.h) or 0)

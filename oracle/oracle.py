"""ctypes front-end of the CPU oracle (oracle/ba_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under megba_b200/ imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libba_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    if force or not os.path.exists(_LIB_PATH) or any(
            os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB_PATH)
            for f in ("ba_oracle.c", "ba_oracle_impl.h")):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True, env={**os.environ, "CC": ""})
    return _LIB_PATH


class LMTrace(C.Structure):
    _fields_ = [("cost", C.c_double), ("accepted", C.c_int), ("pcg_iters", C.c_int),
                ("region", C.c_double), ("dx_norm", C.c_double), ("rho", C.c_double)]


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_graph_create.restype = C.c_void_p
        _lib.orc_graph_create.argtypes = [C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]
        _lib.orc_graph_destroy.argtypes = [C.c_void_p]
        _lib.orc_num_threads.restype = C.c_int
        for sfx in ("_f64", "_f32"):
            getattr(_lib, "orc_cost" + sfx).restype = C.c_double
            getattr(_lib, "orc_jdx_plus_r_sq" + sfx).restype = C.c_double
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Oracle:
    """One problem graph + dtype; methods mirror the oracle's C functions one-to-one."""

    def __init__(self, nC, nP, cam_idx, pt_idx, obs, dtype=np.float64):
        self.L = lib()
        self.dtype = np.dtype(dtype)
        self.sfx = "_f64" if self.dtype == np.float64 else "_f32"
        self.nC, self.nP, self.nE = int(nC), int(nP), int(len(cam_idx))
        self.cam_idx = np.ascontiguousarray(cam_idx, dtype=np.int32)
        self.pt_idx = np.ascontiguousarray(pt_idx, dtype=np.int32)
        self.obs = np.ascontiguousarray(obs, dtype=self.dtype).reshape(self.nE, 2)
        self.G = C.c_void_p(self.L.orc_graph_create(self.nC, self.nP, self.nE, _p(self.cam_idx), _p(self.pt_idx)))

    def __del__(self):
        try:
            self.L.orc_graph_destroy(self.G)
        except Exception:
            pass

    def _f(self, name):
        return getattr(self.L, name + self.sfx)

    def _a(self, a):
        return np.ascontiguousarray(a, dtype=self.dtype)

    def forward(self, cam, pt, jac=True):
        cam, pt = self._a(cam), self._a(pt)
        r = np.empty((self.nE, 2), self.dtype)
        J = np.empty((self.nE, 2, 12), self.dtype) if jac else None
        self._f("orc_forward")(C.c_int64(self.nE), _p(self.cam_idx), _p(self.pt_idx), _p(self.obs),
                               _p(cam), _p(pt), _p(r), _p(J))
        return r, J

    def cost(self, r):
        return float(self._f("orc_cost")(C.c_int64(self.nE), _p(self._a(r))))

    def build(self, r, J, info=None, explicit=True):
        r, J = self._a(r), self._a(J)
        info = None if info is None else self._a(info)
        Hpp = np.empty((self.nC, 9, 9), self.dtype)
        Hll = np.empty((self.nP, 3, 3), self.dtype)
        W = np.empty((self.nE, 9, 3), self.dtype) if explicit else None
        g = np.empty(9 * self.nC + 3 * self.nP, self.dtype)
        self._f("orc_build")(self.G, _p(r), _p(J), _p(info), _p(Hpp), _p(Hll), _p(W), _p(g))
        return Hpp, Hll, W, g

    def process_diag(self, H, region, recover=False, diag=None):
        H = self._a(H).copy()
        num, dim = H.shape[0], H.shape[1]
        diag = np.empty((num, dim), self.dtype) if diag is None else self._a(diag)
        self._f("orc_process_diag")(C.c_int64(num), C.c_int(dim), C.c_double(region), C.c_int(int(recover)),
                                    _p(H), _p(diag))
        return H, diag

    def block_inverse(self, A):
        A = self._a(A)
        out = np.empty_like(A)
        self._f("orc_block_inverse")(C.c_int64(A.shape[0]), C.c_int(A.shape[1]), _p(A), _p(out))
        return out

    def hlp_mul(self, W, J, xc):
        W = None if W is None else self._a(W)
        J = None if J is None else self._a(J)
        t = np.empty(3 * self.nP, self.dtype)
        self._f("orc_hlp_mul")(self.G, _p(W), _p(J), _p(self._a(xc)), _p(t))
        return t

    def hpl_mul(self, W, J, yp):
        W = None if W is None else self._a(W)
        J = None if J is None else self._a(J)
        q = np.empty(9 * self.nC, self.dtype)
        self._f("orc_hpl_mul")(self.G, _p(W), _p(J), _p(self._a(yp)), _p(q))
        return q

    def schur_mul(self, HppD, HllInv, W, J, p):
        W = None if W is None else self._a(W)
        J = None if J is None else self._a(J)
        tmp = np.empty(3 * self.nP, self.dtype)
        q = np.empty(9 * self.nC, self.dtype)
        self._f("orc_schur_mul")(self.G, _p(self._a(HppD)), _p(self._a(HllInv)), _p(W), _p(J),
                                 _p(self._a(p)), _p(tmp), _p(q))
        return q

    def schur_pcg(self, HppD, HllD, W, J, g, dx0, max_iter, tol, refuse_ratio):
        W = None if W is None else self._a(W)
        J = None if J is None else self._a(J)
        dx = self._a(dx0).copy()
        it, rho = C.c_int(0), C.c_double(0)
        done = self._f("orc_schur_pcg")(self.G, _p(self._a(HppD)), _p(self._a(HllD)), _p(W), _p(J),
                                        _p(self._a(g)), C.c_int(max_iter), C.c_double(tol),
                                        C.c_double(refuse_ratio), _p(dx), C.byref(it), C.byref(rho))
        return dx, int(it.value), bool(done), float(rho.value)

    def jdx_plus_r_sq(self, J, r, dx):
        return float(self._f("orc_jdx_plus_r_sq")(self.G, _p(self._a(J)), _p(self._a(r)), _p(self._a(dx))))

    def lm_solve(self, x0, implicit=False, lm_max_iter=20, initial_region=1e3, eps1=1.0, eps2=1e-10,
                 pcg_max_iter=100, pcg_tol=1e-1, pcg_refuse=1.0):
        x = self._a(x0).copy()
        trace = (LMTrace * (lm_max_iter + 2))()
        nt = C.c_int(0)
        k = self._f("orc_lm_solve")(self.G, _p(self.obs), _p(x), C.c_int(int(implicit)), C.c_int(lm_max_iter),
                                    C.c_double(initial_region), C.c_double(eps1), C.c_double(eps2),
                                    C.c_int(pcg_max_iter), C.c_double(pcg_tol), C.c_double(pcg_refuse),
                                    trace, C.byref(nt))
        rows = [dict(cost=t.cost, accepted=t.accepted, pcg_iters=t.pcg_iters, region=t.region,
                     dx_norm=t.dx_norm, rho=t.rho) for t in trace[:nt.value]]
        return x, int(k), rows

    @staticmethod
    def num_threads():
        return int(lib().orc_num_threads())

"""ctypes binding of oracle/liboracle.so (see oracle_bal.h) -- TEST INFRASTRUCTURE."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class LMOptions(C.Structure):
    _fields_ = [("initial_damping", C.c_double), ("iterations", C.c_int64), ("pcg_iterations", C.c_int64),
                ("pcg_tolerance", C.c_double), ("rejection_ratio", C.c_double), ("use_identity", C.c_int),
                ("solver", C.c_int), ("threads", C.c_int)]


def default_options(**kw) -> LMOptions:
    o = LMOptions(1e-4, 50, 10, 1.0, 5.0, 0, 0, 0)
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def build() -> str:
    subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])
    return os.path.join(_HERE, "liboracle.so")


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        dp, ip, lp = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_int64)
        for name in ("orc_create_f64", "orc_create_f32"):
            f = getattr(L, name)
            f.restype = C.c_void_p
            f.argtypes = [C.c_int64, C.c_int64, C.c_int64, ip, ip, dp, dp, dp]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_set_threads.argtypes = [C.c_void_p, C.c_int]
        L.orc_get_params.argtypes = [C.c_void_p, dp, dp]
        L.orc_set_params.argtypes = [C.c_void_p, dp, dp]
        L.orc_set_robust.argtypes = [C.c_void_p, C.c_int, C.c_double, dp]
        L.orc_residuals.restype = C.c_double
        L.orc_residuals.argtypes = [C.c_void_p, dp]
        L.orc_jacobians.argtypes = [C.c_void_p, dp, dp]
        L.orc_linearize.restype = C.c_double
        L.orc_linearize.argtypes = [C.c_void_p, dp, dp]
        L.orc_hessian_structure.argtypes = [C.c_void_p, lp, lp, lp]
        L.orc_hessian_num_values.restype = C.c_int64
        L.orc_hessian_num_values.argtypes = [C.c_void_p]
        L.orc_hessian_values.argtypes = [C.c_void_p, dp]
        L.orc_schur.argtypes = [C.c_void_p, C.c_double, C.c_int, dp, dp]
        L.orc_schur_nnz_blocks.restype = C.c_int64
        L.orc_schur_nnz_blocks.argtypes = [C.c_void_p]
        L.orc_solve.restype = C.c_int64
        L.orc_solve.argtypes = [C.c_void_p, C.POINTER(LMOptions), C.c_double, dp]
        L.orc_lm.restype = C.c_int64
        L.orc_lm.argtypes = [C.c_void_p, C.POINTER(LMOptions), dp]
        L.orc_last_timings.argtypes = [C.c_void_p, dp]
        L.orc_lm_begin.argtypes = [C.c_void_p, C.POINTER(LMOptions)]
        L.orc_lm_step.restype = C.c_int
        L.orc_lm_step.argtypes = [C.c_void_p, C.POINTER(LMOptions), dp]
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class Oracle:
    """CPU restatement of the reference path on one BAL problem."""

    def __init__(self, prob, precision: str = "f64", threads: int = 0):
        L = lib()
        self.L = L
        self.nc, self.np_, self.m = prob.shape()
        self.dimc = 9 * self.nc
        self.dimH = 9 * self.nc + 3 * self.np_
        ci = np.ascontiguousarray(prob.cam_idx, dtype=np.int32)
        pi = np.ascontiguousarray(prob.pt_idx, dtype=np.int32)
        obs = np.ascontiguousarray(prob.obs, dtype=np.float64)
        cams = np.ascontiguousarray(prob.cams, dtype=np.float64)
        pts = np.ascontiguousarray(prob.pts, dtype=np.float64)
        ip = C.POINTER(C.c_int32)
        create = L.orc_create_f64 if precision == "f64" else L.orc_create_f32
        self.h = create(self.nc, self.np_, self.m, ci.ctypes.data_as(ip), pi.ctypes.data_as(ip), _dp(obs), _dp(cams), _dp(pts))
        if threads:
            L.orc_set_threads(self.h, threads)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_destroy(self.h)
            self.h = None

    def params(self):
        c = np.empty((self.nc, 9)); p = np.empty((self.np_, 3))
        self.L.orc_get_params(self.h, _dp(c), _dp(p))
        return c, p

    def set_params(self, cams, pts):
        c = np.ascontiguousarray(cams, dtype=np.float64); p = np.ascontiguousarray(pts, dtype=np.float64)
        self.L.orc_set_params(self.h, _dp(c), _dp(p))

    def set_robust(self, loss="default", delta=0.0, precision=None):
        """Loss (loss.hpp) and per-factor 2x2 precision matrices [m][2][2] (factor.hpp:373-412)."""
        kind = {"default": 0, "huber": 1}[loss]
        P = None if precision is None else np.ascontiguousarray(precision, dtype=np.float64).reshape(-1)
        lib().orc_set_robust(self.h, kind, float(delta), _dp(P) if P is not None else None)

    def residuals(self):
        r = np.empty((self.m, 2))
        chi2 = self.L.orc_residuals(self.h, _dp(r))
        return r, chi2

    def jacobians(self):
        jc = np.empty((self.m, 18)); jp = np.empty((self.m, 6))
        self.L.orc_jacobians(self.h, _dp(jc), _dp(jp))
        return jc, jp

    def linearize(self):
        s = np.empty(self.dimH); b = np.empty(self.dimH)
        chi2 = self.L.orc_linearize(self.h, _dp(s), _dp(b))
        return chi2, s, b

    def hessian_structure(self):
        nblk = self.nc + self.np_
        nnz = self.nc + self.m + self.np_
        cp = np.empty(nblk + 1, dtype=np.int64); ri = np.empty(nnz, dtype=np.int64); off = np.empty(nnz, dtype=np.int64)
        lp = C.POINTER(C.c_int64)
        self.L.orc_hessian_structure(self.h, cp.ctypes.data_as(lp), ri.ctypes.data_as(lp), off.ctypes.data_as(lp))
        return cp, ri, off

    def hessian_values(self):
        v = np.empty(self.L.orc_hessian_num_values(self.h))
        self.L.orc_hessian_values(self.h, _dp(v))
        return v

    def schur(self, mu, use_identity=False, dense=True):
        S = np.zeros((self.dimc, self.dimc), order="F") if dense else None
        bS = np.empty(self.dimc)
        self.L.orc_schur(self.h, float(mu), int(use_identity), _dp(S) if dense else None, _dp(bS))
        return S, bS

    def schur_nnz_blocks(self):
        return int(self.L.orc_schur_nnz_blocks(self.h))

    def solve(self, mu, opts=None):
        opts = opts or default_options()
        d = np.empty(self.dimH)
        k = self.L.orc_solve(self.h, C.byref(opts), float(mu), _dp(d))
        return d, int(k)

    def lm(self, opts=None):
        opts = opts or default_options()
        traj = np.zeros((opts.iterations, 4))
        n = self.L.orc_lm(self.h, C.byref(opts), _dp(traj))
        return traj[:n]

    def lm_begin(self, opts=None):
        self._opts = opts or default_options()
        self.L.orc_lm_begin(self.h, C.byref(self._opts))

    def lm_step(self):
        out = np.zeros(4)
        go = self.L.orc_lm_step(self.h, C.byref(self._opts), _dp(out))
        return out, bool(go)

    def timings(self):
        t = np.zeros(6)
        self.L.orc_last_timings(self.h, _dp(t))
        return dict(zip(["linearize", "hessian", "schur", "solve", "backsubst", "update_cost"], t.tolist()))

"""ctypes binding of the generic factor-graph C ABI (``include/graphite_b200_graph.h``).

Harness plumbing for the tests and examples; the product is the shared library.  Factor evaluation is the CALLER's CUDA
code: callbacks are C function pointers (e.g. from a user ``.so`` compiled with nvcc), never Python.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import binding
from .binding import LMOptions, LMResult, PcgOptions, SolveInfo, _DT, _NP, _ptr

GB_MAX_ARITY = 4

# every symbol include/graphite_b200_graph.h declares
SYMBOLS = [
    "gb_graph_create", "gb_graph_destroy", "gb_graph_add_vertex_set", "gb_graph_add_factor_set", "gb_graph_set_update",
    "gb_graph_set_fixed", "gb_graph_set_active", "gb_graph_set_vertices", "gb_graph_get_vertices", "gb_graph_vertices_device",
    "gb_graph_set_precision", "gb_graph_set_loss", "gb_graph_set_scaling", "gb_graph_initialize", "gb_graph_vertex_columns",
    "gb_graph_hessian_structure", "gb_graph_linearize", "gb_graph_cost", "gb_graph_get", "gb_graph_hessian_values",
    "gb_graph_jv", "gb_graph_jtpv", "gb_graph_set_damping", "gb_graph_solve", "gb_graph_lm",
    "gb_graph_bind_linearization", "gb_graph_bind_gradient", "gb_graph_update_values", "gb_graph_solve_device",
]


class VertexSetDesc(C.Structure):
    _fields_ = [("dimension", C.c_int32), ("parameters", C.c_int32), ("count", C.c_int64), ("global_ids", C.c_void_p),
                ("fixed", C.c_void_p), ("eliminate", C.c_int32), ("reserved", C.c_int32)]


class FactorSetDesc(C.Structure):
    _fields_ = [("residual_dim", C.c_int32), ("arity", C.c_int32), ("vertex_set", C.c_int32 * GB_MAX_ARITY),
                ("count", C.c_int64), ("vertex_index", C.c_void_p), ("active", C.c_void_p), ("loss", C.c_int32),
                ("reserved", C.c_int32), ("loss_delta", C.c_double)]


def _bind(L):
    if getattr(L, "_graph_bound", False):
        return
    vp = C.c_void_p
    L.gb_graph_create.argtypes = [vp, C.c_int, C.c_int, C.POINTER(vp)]
    L.gb_graph_destroy.argtypes = [vp]
    L.gb_graph_add_vertex_set.argtypes = [vp, C.POINTER(VertexSetDesc)]
    L.gb_graph_add_factor_set.argtypes = [vp, C.POINTER(FactorSetDesc), vp, vp]
    L.gb_graph_set_update.argtypes = [vp, C.c_int, vp, vp]
    L.gb_graph_set_fixed.argtypes = [vp, C.c_int, vp]
    L.gb_graph_set_active.argtypes = [vp, C.c_int, vp]
    L.gb_graph_set_vertices.argtypes = [vp, C.c_int, vp]
    L.gb_graph_get_vertices.argtypes = [vp, C.c_int, vp]
    L.gb_graph_vertices_device.argtypes = [vp, C.c_int, C.POINTER(vp)]
    L.gb_graph_set_precision.argtypes = [vp, C.c_int, vp]
    L.gb_graph_set_loss.argtypes = [vp, C.c_int, C.c_int, C.c_double]
    L.gb_graph_set_scaling.argtypes = [vp, C.c_int]
    L.gb_graph_initialize.argtypes = [vp, C.c_int, C.POINTER(C.c_int64)]
    L.gb_graph_vertex_columns.argtypes = [vp, C.c_int, vp]
    L.gb_graph_hessian_structure.argtypes = [vp, vp, vp, vp]
    L.gb_graph_linearize.argtypes = [vp, C.POINTER(C.c_double)]
    L.gb_graph_cost.argtypes = [vp, C.POINTER(C.c_double)]
    L.gb_graph_get.argtypes = [vp, C.c_int, C.c_int, vp]
    L.gb_graph_hessian_values.argtypes = [vp, vp]
    L.gb_graph_jv.argtypes = [vp, vp, vp]
    L.gb_graph_jtpv.argtypes = [vp, vp, vp]
    L.gb_graph_set_damping.argtypes = [vp, C.c_double, C.c_int]
    L.gb_graph_solve.argtypes = [vp, C.POINTER(PcgOptions), vp, C.POINTER(SolveInfo)]
    L.gb_graph_lm.argtypes = [vp, C.POINTER(LMOptions), C.POINTER(LMResult), vp]
    L._graph_bound = True


class Graph:
    """A generic factor graph on one GPU."""

    def __init__(self, ctx: binding.Context, precision: str = "f64-f64"):
        self.ctx, self.L = ctx, ctx.L
        _bind(self.L)
        t, s = precision.split("-")
        self.T = _NP[t]
        self.S = _NP[s] if s != "bf16" else _NP[t]  # bf16 values are exported as T
        h = C.c_void_p()
        ctx.check(self.L.gb_graph_create(ctx.h, _DT[t], _DT[s], C.byref(h)))
        self.h = h
        self.vsets, self.fsets = [], []  # (dim, npar, count) / (E, arity, sets, count)
        self.info = None

    def close(self):
        if getattr(self, "h", None):
            if self.ctx.h:  # a graph that outlived its context (garbage collection order) has nothing left to free
                self.L.gb_graph_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _id(self, rc):
        if rc < 0:
            self.ctx.check(rc)
        return rc

    def add_vertex_set(self, dimension, global_ids, fixed=None, eliminate=False, parameters=0):
        gid = np.ascontiguousarray(global_ids, dtype=np.int64)
        fx = None if fixed is None else np.ascontiguousarray(fixed, dtype=np.uint8)
        d = VertexSetDesc(dimension, parameters, len(gid), _ptr(gid), None if fx is None else _ptr(fx), 1 if eliminate else 0, 0)
        i = self._id(self.L.gb_graph_add_vertex_set(self.h, C.byref(d)))
        self.vsets.append((dimension, parameters or dimension, len(gid)))
        return i

    def add_factor_set(self, residual_dim, vertex_sets, vertex_index, fn, user=None, active=None, loss=0, loss_delta=0.0):
        vi = np.ascontiguousarray(vertex_index, dtype=np.int32).reshape(-1, len(vertex_sets))
        ac = None if active is None else np.ascontiguousarray(active, dtype=np.uint8)
        vs = (C.c_int32 * GB_MAX_ARITY)(*(list(vertex_sets) + [0] * (GB_MAX_ARITY - len(vertex_sets))))
        d = FactorSetDesc(residual_dim, len(vertex_sets), vs, vi.shape[0], _ptr(vi), None if ac is None else _ptr(ac), loss, 0, loss_delta)
        i = self._id(self.L.gb_graph_add_factor_set(self.h, C.byref(d), fn, user))
        self.fsets.append((residual_dim, len(vertex_sets), list(vertex_sets), vi.shape[0]))
        return i

    def set_update(self, vset, fn, user=None):
        self.ctx.check(self.L.gb_graph_set_update(self.h, vset, fn, user))

    def set_fixed(self, vset, fixed):
        fx = np.ascontiguousarray(fixed, dtype=np.uint8)
        self.ctx.check(self.L.gb_graph_set_fixed(self.h, vset, _ptr(fx)))

    def set_active(self, fset, active):
        a = np.ascontiguousarray(active, dtype=np.uint8)
        self.ctx.check(self.L.gb_graph_set_active(self.h, fset, _ptr(a)))

    def set_vertices(self, vset, values):
        v = np.ascontiguousarray(values, dtype=self.T)
        assert v.size == self.vsets[vset][1] * self.vsets[vset][2]
        self.ctx.check(self.L.gb_graph_set_vertices(self.h, vset, _ptr(v)))

    def get_vertices(self, vset):
        d, npar, n = self.vsets[vset]
        out = np.empty((n, npar), dtype=self.T)
        self.ctx.check(self.L.gb_graph_get_vertices(self.h, vset, _ptr(out)))
        return out

    def set_precision(self, fset, P):
        if P is None:
            self.ctx.check(self.L.gb_graph_set_precision(self.h, fset, None))
            return
        p = np.ascontiguousarray(P, dtype=self.T)
        self.ctx.check(self.L.gb_graph_set_precision(self.h, fset, _ptr(p)))

    def set_loss(self, fset, loss, delta=0.0):
        self.ctx.check(self.L.gb_graph_set_loss(self.h, fset, loss, delta))

    def set_scaling(self, on):
        self.ctx.check(self.L.gb_graph_set_scaling(self.h, 1 if on else 0))

    def initialize(self, level=0):
        info = (C.c_int64 * 8)()
        self.ctx.check(self.L.gb_graph_initialize(self.h, level, info))
        keys = ["hessian_dim", "block_columns", "hessian_blocks", "hessian_values", "residual_rows", "active_factors", "device_bytes"]
        self.info = dict(zip(keys, [int(x) for x in info]))
        return self.info

    def vertex_columns(self, vset):
        out = np.empty(self.vsets[vset][2], dtype=np.int64)
        self.ctx.check(self.L.gb_graph_vertex_columns(self.h, vset, _ptr(out)))
        return out

    def hessian_structure(self):
        cp = np.empty(self.info["block_columns"] + 1, dtype=np.int64)
        ri = np.empty(self.info["hessian_blocks"], dtype=np.int64)
        off = np.empty(self.info["hessian_blocks"], dtype=np.int64)
        self.ctx.check(self.L.gb_graph_hessian_structure(self.h, _ptr(cp), _ptr(ri), _ptr(off)))
        return cp, ri, off

    def linearize(self):
        c = C.c_double()
        self.ctx.check(self.L.gb_graph_linearize(self.h, C.byref(c)))
        return c.value

    def cost(self):
        c = C.c_double()
        self.ctx.check(self.L.gb_graph_cost(self.h, C.byref(c)))
        return c.value

    def _get(self, which, set_, shape):
        out = np.empty(shape, dtype=self.T)
        self.ctx.check(self.L.gb_graph_get(self.h, which, set_, _ptr(out)))
        return out

    def gradient(self):
        return self._get(0, 0, self.info["hessian_dim"])

    def scales(self):
        return self._get(1, 0, self.info["hessian_dim"])

    def scalar_diagonal(self):
        return self._get(17, 0, self.info["hessian_dim"])

    def residuals(self, fset):
        E, _, _, n = self.fsets[fset]
        return self._get(2, fset, (n, E))

    def chi2_per_factor(self, fset):
        return self._get(3, fset, self.fsets[fset][3])

    def loss_derivative(self, fset):
        return self._get(4, fset, self.fsets[fset][3])

    def jacobians(self, fset, slot):
        E, _, sets, n = self.fsets[fset]
        d = self.vsets[sets[slot]][0]
        return self._get(5 + slot, fset, (n, d, E))  # [factor][column][row]: column-major E x d

    def block_diagonal(self, vset):
        d, _, n = self.vsets[vset]
        return self._get(16, vset, (n, d, d))  # [vertex][col][row]

    def hessian_values(self):
        out = np.empty(self.info["hessian_values"], dtype=self.S)
        self.ctx.check(self.L.gb_graph_hessian_values(self.h, _ptr(out)))
        return out

    def jv(self, x):
        xx = np.ascontiguousarray(x, dtype=self.T)
        y = np.empty(self.info["residual_rows"], dtype=self.T)
        self.ctx.check(self.L.gb_graph_jv(self.h, _ptr(xx), _ptr(y)))
        return y

    def jtpv(self, v):
        vv = np.ascontiguousarray(v, dtype=self.T)
        y = np.empty(self.info["hessian_dim"], dtype=self.T)
        self.ctx.check(self.L.gb_graph_jtpv(self.h, _ptr(vv), _ptr(y)))
        return y

    def set_damping(self, mu, use_identity=False):
        self.ctx.check(self.L.gb_graph_set_damping(self.h, mu, 1 if use_identity else 0))

    def solve(self, max_iterations=10, tolerance=1.0, rejection_ratio=5.0):
        o = PcgOptions(max_iterations, tolerance, rejection_ratio, 1, 0)
        info = SolveInfo()
        x = np.empty(self.info["hessian_dim"], dtype=self.T)
        self.ctx.check(self.L.gb_graph_solve(self.h, C.byref(o), _ptr(x), C.byref(info)))
        return x, {"pcg_iterations": int(info.pcg_iterations), "rz_final": info.rz_final, "stop_reason": int(info.stop_reason)}

    def lm(self, iterations=50, initial_damping=1e-4, use_identity=False, pcg_iterations=10, pcg_tolerance=1.0,
           rejection_ratio=5.0, verbose=False, early_stop=False):
        o = LMOptions()
        o.initial_damping, o.iterations, o.use_identity, o.verbose = initial_damping, iterations, 1 if use_identity else 0, 1 if verbose else 0
        o.pcg = PcgOptions(pcg_iterations, pcg_tolerance, rejection_ratio, 1, 0)
        o.early_stop = 1 if early_stop else 0
        res = LMResult()
        traj = np.zeros((max(iterations, 1), 4), dtype=np.float64)
        self.ctx.check(self.L.gb_graph_lm(self.h, C.byref(o), C.byref(res), _ptr(traj)))
        keys = ["iterations", "initial_chi2", "final_chi2", "final_damping", "accepted", "rejected", "pcg_iterations_total",
                "seconds_total", "final_nu", "termination"]
        return traj[: res.iterations], {k: getattr(res, k) for k in keys}

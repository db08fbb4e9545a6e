"""ctypes binding of ``libgraphite_b200.so`` (C ABI in ``include/graphite_b200.h``).

This is harness plumbing for tests and ``bench.py``; the product is the shared library.  There is
no fallback: if the library is missing or no B200 is visible, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GRAPHITE_B200_LIB") or os.path.join(_HERE, "libgraphite_b200.so")  # env: A/B builds of the same ABI

GB_F32, GB_F64 = 0, 1
_DT = {"f32": GB_F32, "f64": GB_F64, "bf16": 2}
_NP = {"f32": np.float32, "f64": np.float64, "bf16": np.uint16}  # (bf16: storage precision only, never exported)

# every symbol include/graphite_b200.h declares
SYMBOLS = [
    "gb_version", "gb_context_create", "gb_context_destroy", "gb_last_error", "gb_comm_unique_id", "gb_comm_init",
    "gb_problem_create", "gb_problem_destroy", "gb_problem_info", "gb_set_observations", "gb_stage_observations_async", "gb_commit_observations", "gb_set_vertices",
    "gb_get_vertices", "gb_set_factor", "gb_set_loss", "gb_set_precision", "gb_hessian_structure", "gb_linearize", "gb_compute_cost", "gb_get_gradient", "gb_get_scales",
    "gb_get_residuals", "gb_get_jacobians", "gb_hessian_values", "gb_set_damping", "gb_solve", "gb_get_schur_rhs",
    "gb_get_schur_diagonal", "gb_schur_multiply", "gb_schur_structure", "gb_schur_values", "gb_try_step", "gb_revert_step", "gb_lm", "gb_kernel_launches",
    "gb_time_stage", "gb_structure_create", "gb_structure_destroy", "gb_structure_info", "gb_structure_array",
    "gb_structure_hessian", "gb_structure_schur", "gb_context_create_on_stream", "gb_set_observations_device",
    "gb_set_vertices_device", "gb_get_vertices_device", "gb_import_linearization", "gb_solve_device", "gb_schur_csc", "gb_set_fixed", "gb_problem_structure_array",
]


class ProblemDesc(C.Structure):
    _fields_ = [("precision_T", C.c_int32), ("precision_S", C.c_int32), ("num_cameras", C.c_int64),
                ("num_points", C.c_int64), ("num_observations", C.c_int64), ("camera_index", C.POINTER(C.c_int32)),
                ("point_index", C.POINTER(C.c_int32)), ("tile_size", C.c_int32), ("slot_cap", C.c_int32),
                ("super_tile_observations", C.c_int64), ("flags", C.c_int64)]


class PcgOptions(C.Structure):
    _fields_ = [("max_iterations", C.c_int64), ("tolerance", C.c_double), ("rejection_ratio", C.c_double),
                ("solver", C.c_int32), ("schur_mode", C.c_int32)]


class SolveInfo(C.Structure):
    _fields_ = [("pcg_iterations", C.c_int64), ("rz_final", C.c_double), ("stop_reason", C.c_int32), ("schur_mode", C.c_int32)]


class LMOptions(C.Structure):
    _fields_ = [("initial_damping", C.c_double), ("iterations", C.c_int64), ("use_identity", C.c_int32),
                ("verbose", C.c_int32), ("pcg", PcgOptions), ("stop_flag", C.POINTER(C.c_int32)),
                ("resume", C.c_int32), ("profile_product", C.c_int32), ("initial_nu", C.c_double),
                ("defer_final_linearize", C.c_int32), ("early_stop", C.c_int32)]


class LMResult(C.Structure):
    _fields_ = [("iterations", C.c_int64), ("initial_chi2", C.c_double), ("final_chi2", C.c_double),
                ("final_damping", C.c_double), ("accepted", C.c_int64), ("rejected", C.c_int64),
                ("pcg_iterations_total", C.c_int64), ("seconds_total", C.c_double), ("seconds_linearize", C.c_double),
                ("seconds_prepare", C.c_double), ("seconds_pcg", C.c_double), ("seconds_backsubst", C.c_double),
                ("seconds_cost", C.c_double), ("final_nu", C.c_double), ("product_launches", C.c_int64),
                ("product_seconds", C.c_double), ("update_seconds", C.c_double), ("termination", C.c_int32),
                ("reserved", C.c_int32), ("pcg_phase_seconds", C.c_double * 6)]


SOLVERS = {"pcg-schur": 0, "pcg": 1, "direct-schur": 2}  # names of examples/bal.cu --solver (eigen-schur / cudss-schur: direct)
SCHUR_MODES = {"auto": 0, "implicit": 1, "explicit": 2}

_lib = None


class GraphiteB200Error(RuntimeError):
    pass


def load_library():
    """Load the CUDA library; raises if it has not been built (no silent fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GraphiteB200Error(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.gb_version.restype = C.c_int
    L.gb_context_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.gb_context_destroy.argtypes = [vp]
    L.gb_last_error.restype = C.c_char_p
    L.gb_last_error.argtypes = [vp]
    L.gb_comm_unique_id.argtypes = [vp]
    L.gb_comm_init.argtypes = [vp, C.c_int, C.c_int, vp]
    L.gb_problem_create.argtypes = [vp, C.POINTER(ProblemDesc), C.POINTER(vp)]
    L.gb_problem_destroy.argtypes = [vp]
    L.gb_problem_info.argtypes = [vp, C.POINTER(C.c_int64)]
    L.gb_set_observations.argtypes = [vp, vp]
    L.gb_set_vertices.argtypes = [vp, vp, vp]
    L.gb_stage_observations_async.argtypes = [vp, vp, C.c_int]
    L.gb_commit_observations.argtypes = [vp, C.c_int]
    L.gb_get_vertices.argtypes = [vp, vp, vp]
    L.gb_set_factor.argtypes = [vp, vp, vp]
    L.gb_set_loss.argtypes = [vp, C.c_int, C.c_double]
    L.gb_set_fixed.argtypes = [vp, vp, vp]
    L.gb_set_precision.argtypes = [vp, vp]
    L.gb_hessian_structure.argtypes = [vp, vp, vp, vp]
    L.gb_linearize.argtypes = [vp, C.POINTER(C.c_double)]
    L.gb_compute_cost.argtypes = [vp, C.POINTER(C.c_double)]
    for n in ("gb_get_gradient", "gb_get_scales", "gb_get_residuals", "gb_hessian_values", "gb_get_schur_rhs",
              "gb_get_schur_diagonal"):
        getattr(L, n).argtypes = [vp, vp]
    L.gb_get_jacobians.argtypes = [vp, vp, vp]
    L.gb_set_damping.argtypes = [vp, C.c_double, C.c_int]
    L.gb_solve.argtypes = [vp, C.POINTER(PcgOptions), vp, C.POINTER(SolveInfo)]
    L.gb_schur_multiply.argtypes = [vp, vp, vp]
    L.gb_schur_structure.argtypes = [vp, vp, vp, C.POINTER(C.c_int64)]
    L.gb_schur_values.argtypes = [vp, vp]
    L.gb_schur_csc.argtypes = [vp, vp, vp, vp, C.POINTER(C.c_int64)]
    L.gb_try_step.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.gb_revert_step.argtypes = [vp]
    L.gb_lm.argtypes = [vp, C.POINTER(LMOptions), C.POINTER(LMResult), vp]
    L.gb_kernel_launches.restype = C.c_int64
    L.gb_kernel_launches.argtypes = [vp]
    L.gb_time_stage.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_double)]
    L.gb_structure_create.argtypes = [C.POINTER(ProblemDesc), C.POINTER(vp), C.c_char_p, C.c_int]
    L.gb_structure_destroy.argtypes = [vp]
    L.gb_structure_info.argtypes = [vp, C.POINTER(C.c_int64)]
    L.gb_structure_array.argtypes = [vp, C.c_int, vp, C.POINTER(C.c_int64)]
    L.gb_problem_structure_array.argtypes = [vp, C.c_int, vp, C.POINTER(C.c_int64)]
    L.gb_structure_hessian.argtypes = [vp, vp, vp, vp]
    L.gb_structure_schur.argtypes = [vp, vp, vp, C.POINTER(C.c_int64)]
    _lib = L
    return L


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Context:
    def __init__(self, device: int = 0):
        self.L = load_library()
        h = C.c_void_p()
        rc = self.L.gb_context_create(device, C.byref(h))
        if rc != 0:
            raise GraphiteB200Error(f"gb_context_create(device={device}) failed with {rc}: a B200 (sm_100) GPU is required")
        self.h = h
        self.device = device

    def check(self, rc):
        if rc != 0:
            raise GraphiteB200Error(f"graphite_b200 error {rc}: {self.L.gb_last_error(self.h).decode()}")

    def comm_init(self, nranks: int, rank: int, unique_id: bytes):
        buf = C.create_string_buffer(unique_id, 128)
        self.check(self.L.gb_comm_init(self.h, nranks, rank, buf))

    @staticmethod
    def comm_unique_id() -> bytes:
        L = load_library()
        buf = C.create_string_buffer(128)
        rc = L.gb_comm_unique_id(buf)
        if rc != 0:
            raise GraphiteB200Error(f"gb_comm_unique_id failed with {rc}")
        return buf.raw

    def kernel_launches(self) -> int:
        return int(self.L.gb_kernel_launches(self.h))

    def close(self):
        if self.h:
            self.L.gb_context_destroy(self.h)
            self.h = None


class Problem:
    """One BAL problem on one GPU (one rank's point partition)."""

    def __init__(self, ctx: Context, cam_idx, pt_idx, n_cams: int, n_pts: int, precision: str = "f64-f64", tile_size: int = 0,
                 slot_cap: int = 0, super_tile_observations: int = 0, partition: bool = False, host_tables: bool = False):
        self.ctx, self.L = ctx, ctx.L
        t, s = precision.split("-")
        self.T, self.S = _NP[t], _NP[s]
        self.precision = precision
        self.n_cams, self.n_pts, self.n_obs = int(n_cams), int(n_pts), int(len(cam_idx))
        self.dimc = 9 * self.n_cams
        self.dimH = 9 * self.n_cams + 3 * self.n_pts
        ci = np.ascontiguousarray(cam_idx, dtype=np.int32)
        pi = np.ascontiguousarray(pt_idx, dtype=np.int32)
        d = ProblemDesc(_DT[t], _DT[s], self.n_cams, self.n_pts, self.n_obs, ci.ctypes.data_as(C.POINTER(C.c_int32)),
                        pi.ctypes.data_as(C.POINTER(C.c_int32)), tile_size, slot_cap, super_tile_observations,
                        (1 if partition else 0) | (2 if host_tables else 0))
        h = C.c_void_p()
        ctx.check(self.L.gb_problem_create(ctx.h, C.byref(d), C.byref(h)))
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.L.gb_problem_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def structure_array(self, which: int):
        """A structure table as it is on the device (10 slot_of_obs, 13 ometa, 17 tile records, 18 tile_cam, 19 cm_slot, 20 cm_pt)."""
        n = C.c_int64()
        self.ctx.check(self.L.gb_problem_structure_array(self.h, which, None, C.byref(n)))
        out = np.empty(n.value, dtype=np.uint8 if which == 17 else (np.uint32 if which == 13 else np.int32))
        self.ctx.check(self.L.gb_problem_structure_array(self.h, which, _ptr(out), C.byref(n)))
        return out

    def info(self):
        a = (C.c_int64 * 12)()
        self.ctx.check(self.L.gb_problem_info(self.h, a))
        return dict(zip(INFO_KEYS, [int(v) for v in a]))

    # ---- data ------------------------------------------------------------------------------------
    def set_observations(self, obs):
        o = np.ascontiguousarray(obs, dtype=self.T)
        assert o.shape == (self.n_obs, 2)
        self.ctx.check(self.L.gb_set_observations(self.h, _ptr(o)))

    def set_vertices(self, cams, pts):
        c = np.ascontiguousarray(cams, dtype=self.T)
        p = np.ascontiguousarray(pts, dtype=self.T)
        assert c.shape == (self.n_cams, 9) and p.shape == (self.n_pts, 3)
        self.ctx.check(self.L.gb_set_vertices(self.h, _ptr(c), _ptr(p)))

    def set_factor(self, fn_ptr, user_ptr=None):
        """User-defined factor: fn_ptr = address of a `gb_factor_fn` (int), or None for the built-in BAL factor."""
        self.ctx.check(self.L.gb_set_factor(self.h, C.c_void_p(fn_ptr) if fn_ptr else None,
                                            C.c_void_p(user_ptr) if user_ptr else None))

    def set_fixed(self, cameras_fixed=None, points_fixed=None):
        """VertexDescriptor::set_fixed for cameras / points (uint8 masks, None = none fixed)."""
        fc = None if cameras_fixed is None else np.ascontiguousarray(cameras_fixed, dtype=np.uint8)
        fp = None if points_fixed is None else np.ascontiguousarray(points_fixed, dtype=np.uint8)
        assert fc is None or fc.size == self.n_cams
        assert fp is None or fp.size == self.n_pts
        self.ctx.check(self.L.gb_set_fixed(self.h, None if fc is None else _ptr(fc), None if fp is None else _ptr(fp)))

    def set_loss(self, loss: str = "default", delta: float = 0.0):
        self.ctx.check(self.L.gb_set_loss(self.h, {"default": 0, "huber": 1}[loss], float(delta)))

    def set_precision(self, precision):
        """[n_obs][2][2] SPD matrices in the caller's factor order, or None for the identity."""
        if precision is None:
            self.ctx.check(self.L.gb_set_precision(self.h, None))
            return
        P = np.ascontiguousarray(precision, dtype=self.T).reshape(-1)
        assert P.size == 4 * self.n_obs
        self.ctx.check(self.L.gb_set_precision(self.h, _ptr(P)))

    def set_vertices_raw(self, cams_ptr: int, pts_ptr: int):
        """Host pointers (e.g. pinned torch tensors) of dtype T, shapes [n_cams,9] / [n_pts,3]."""
        self.ctx.check(self.L.gb_set_vertices(self.h, C.c_void_p(cams_ptr), C.c_void_p(pts_ptr)))

    def set_observations_raw(self, obs_ptr: int):
        self.ctx.check(self.L.gb_set_observations(self.h, C.c_void_p(obs_ptr)))

    def stage_observations_async(self, obs_ptr: int, slot: int):
        """Start the H2D copy of a pinned host buffer ([n_obs][2] of T) into staging slot 0/1; returns at once."""
        self.ctx.check(self.L.gb_stage_observations_async(self.h, C.c_void_p(obs_ptr), slot))

    def commit_observations(self, slot: int):
        self.ctx.check(self.L.gb_commit_observations(self.h, slot))

    def get_vertices(self):
        c = np.empty((self.n_cams, 9), dtype=self.T)
        p = np.empty((self.n_pts, 3), dtype=self.T)
        self.ctx.check(self.L.gb_get_vertices(self.h, _ptr(c), _ptr(p)))
        return c, p

    def get_vertices_raw(self, cams_ptr: int, pts_ptr: int):
        self.ctx.check(self.L.gb_get_vertices(self.h, C.c_void_p(cams_ptr), C.c_void_p(pts_ptr)))

    # ---- structure ---------------------------------------------------------------------------------
    def hessian_structure(self):
        nblk = self.n_cams + self.n_pts
        nnz = self.n_cams + self.n_obs + self.n_pts
        cp = np.empty(nblk + 1, dtype=np.int64)
        ri = np.empty(nnz, dtype=np.int64)
        off = np.empty(nnz, dtype=np.int64)
        self.ctx.check(self.L.gb_hessian_structure(self.h, _ptr(cp), _ptr(ri), _ptr(off)))
        return cp, ri, off

    # ---- stages --------------------------------------------------------------------------------------
    def linearize(self) -> float:
        v = C.c_double()
        self.ctx.check(self.L.gb_linearize(self.h, C.byref(v)))
        return v.value

    def compute_cost(self) -> float:
        v = C.c_double()
        self.ctx.check(self.L.gb_compute_cost(self.h, C.byref(v)))
        return v.value

    def _get(self, fn, n, dtype):
        a = np.empty(n, dtype=dtype)
        self.ctx.check(fn(self.h, _ptr(a)))
        return a

    def gradient(self):
        return self._get(self.L.gb_get_gradient, self.dimH, self.T)

    def scales(self):
        return self._get(self.L.gb_get_scales, self.dimH, self.T)

    def residuals(self):
        return self._get(self.L.gb_get_residuals, 2 * self.n_obs, self.T).reshape(self.n_obs, 2)

    def jacobians(self):
        jc = np.empty((self.n_obs, 18))
        jp = np.empty((self.n_obs, 6))
        self.ctx.check(self.L.gb_get_jacobians(self.h, _ptr(jc), _ptr(jp)))
        return jc, jp

    def hessian_values(self):
        n = 81 * self.n_cams + 27 * self.n_obs + 9 * self.n_pts
        return self._get(self.L.gb_hessian_values, n, self.S)

    def set_damping(self, mu: float, use_identity: bool = False):
        self.ctx.check(self.L.gb_set_damping(self.h, float(mu), int(use_identity)))

    def solve(self, max_iterations=10, tolerance=1.0, rejection_ratio=5.0, want_delta=True, solver="pcg-schur", schur_mode="auto"):
        o = PcgOptions(max_iterations, tolerance, rejection_ratio, SOLVERS[solver], SCHUR_MODES[schur_mode])
        info = SolveInfo()
        d = np.empty(self.dimH, dtype=self.T) if want_delta else None
        self.ctx.check(self.L.gb_solve(self.h, C.byref(o), _ptr(d) if want_delta else None, C.byref(info)))
        return d, {"pcg_iterations": int(info.pcg_iterations), "rz_final": info.rz_final, "stop_reason": int(info.stop_reason),
                   "schur_mode": int(info.schur_mode)}

    def schur_rhs(self):
        return self._get(self.L.gb_get_schur_rhs, self.dimc, self.T)

    def schur_diagonal(self):
        return self._get(self.L.gb_get_schur_diagonal, 81 * self.n_cams, self.T).reshape(self.n_cams, 9, 9).transpose(0, 2, 1)

    def schur_multiply(self, x):
        xx = np.ascontiguousarray(x, dtype=self.T)
        y = np.empty(self.dimc, dtype=self.T)
        self.ctx.check(self.L.gb_schur_multiply(self.h, _ptr(xx), _ptr(y)))
        return y

    def schur_structure(self):
        """Upper block-CSC of S: (colptr [n_cams+1], rowidx [nnz])."""
        n = C.c_int64()
        self.ctx.check(self.L.gb_schur_structure(self.h, None, None, C.byref(n)))
        cp = np.empty(self.n_cams + 1, dtype=np.int64)
        ri = np.empty(n.value, dtype=np.int64)
        self.ctx.check(self.L.gb_schur_structure(self.h, _ptr(cp), _ptr(ri), C.byref(n)))
        return cp, ri

    def schur_values(self):
        """Explicit S at the current damping: [nnz][9][9] blocks (row, column) in structure order."""
        cp, ri = self.schur_structure()
        v = np.empty(len(ri) * 81, dtype=self.T)
        self.ctx.check(self.L.gb_schur_values(self.h, _ptr(v)))
        return v.reshape(len(ri), 9, 9).transpose(0, 2, 1)

    def schur_csc(self):
        """S as the reference's scalar upper CSC (csc_utils.hpp:73-193): (pointers, indices, values)."""
        n = C.c_int64()
        self.ctx.check(self.L.gb_schur_csc(self.h, None, None, None, C.byref(n)))
        ptr = np.empty(9 * self.n_cams + 1, dtype=np.int32)
        idx = np.empty(n.value, dtype=np.int32)
        val = np.empty(n.value, dtype=self.T)
        self.ctx.check(self.L.gb_schur_csc(self.h, _ptr(ptr), _ptr(idx), _ptr(val), C.byref(n)))
        return ptr, idx, val

    def try_step(self):
        a, b = C.c_double(), C.c_double()
        self.ctx.check(self.L.gb_try_step(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def revert_step(self):
        self.ctx.check(self.L.gb_revert_step(self.h))

    def lm(self, iterations=50, initial_damping=1e-4, pcg_iterations=10, pcg_tolerance=1.0, rejection_ratio=5.0,
           use_identity=False, verbose=False, resume=False, initial_nu=2.0, profile_product=False, solver="pcg-schur",
           defer_final_linearize=False, early_stop=False, schur_mode="auto"):
        o = LMOptions(initial_damping, iterations, int(use_identity), int(verbose),
                      PcgOptions(pcg_iterations, pcg_tolerance, rejection_ratio, SOLVERS[solver], SCHUR_MODES[schur_mode]), None, int(resume),
                      int(profile_product), float(initial_nu), int(defer_final_linearize), int(early_stop))
        res = LMResult()
        traj = np.zeros((max(iterations, 1), 4))
        self.ctx.check(self.L.gb_lm(self.h, C.byref(o), C.byref(res), _ptr(traj)))
        out = {k: getattr(res, k) for k, _ in LMResult._fields_}
        out["pcg_phase_seconds"] = [float(v) for v in res.pcg_phase_seconds]
        return traj[: res.iterations], out

    def time_stage(self, stage: int, repetitions: int) -> float:
        v = C.c_double()
        self.ctx.check(self.L.gb_time_stage(self.h, stage, repetitions, C.byref(v)))
        return v.value


def problem_from_bal(ctx: Context, prob, precision="f64-f64", tile_size=0, slot_cap=0, super_tile_observations=0,
                     partition=False, host_tables=False) -> Problem:
    p = Problem(ctx, prob.cam_idx, prob.pt_idx, prob.n_cams, prob.n_pts, precision, tile_size, slot_cap,
                super_tile_observations, partition, host_tables)
    p.set_observations(prob.obs)
    p.set_vertices(prob.cams, prob.pts)
    return p


STRUCT_ARRAYS = ["cam_idx", "pt_idx", "pptr", "tile_obs", "tile_pt", "st_tile", "st_row", "row_cam", "cam_row_ptr",
                 "cam_row_list", "slot_of_obs", "rank", "perm", "ometa", "seg_tab", "pt_tab", "tmeta"]
_STRUCT_DTYPES = {"rank": np.uint8, "perm": np.int64, "ometa": np.uint32, "seg_tab": np.uint32, "pt_tab": np.uint16}
INFO_KEYS = ["n_tiles", "n_partial_rows", "max_track", "hessian_dim", "n_hessian_blocks", "n_hessian_values",
             "device_bytes", "n_obs", "n_super_tiles", "n_camera_segments", "storage_slots", "exchange_mode"]


def host_structure(cam_idx, pt_idx, n_cams: int, n_pts: int, tile_size: int = 0, slot_cap: int = 0,
                   super_tile_observations: int = 0):
    """Structure build on the host only (no GPU): dict of arrays + info + Hessian block CSC."""
    L = load_library()
    ci = np.ascontiguousarray(cam_idx, dtype=np.int32)
    pi = np.ascontiguousarray(pt_idx, dtype=np.int32)
    d = ProblemDesc(GB_F64, GB_F64, int(n_cams), int(n_pts), int(len(ci)), ci.ctypes.data_as(C.POINTER(C.c_int32)),
                    pi.ctypes.data_as(C.POINTER(C.c_int32)), tile_size, slot_cap, super_tile_observations, 0)
    h = C.c_void_p()
    err = C.create_string_buffer(256)
    rc = L.gb_structure_create(C.byref(d), C.byref(h), err, 256)
    if rc != 0:
        raise GraphiteB200Error(f"structure: {err.value.decode()} ({rc})")
    try:
        out = {}
        for i, name in enumerate(STRUCT_ARRAYS):
            n = C.c_int64()
            L.gb_structure_array(h, i, None, C.byref(n))
            a = np.empty(n.value, dtype=_STRUCT_DTYPES.get(name, np.int32))
            L.gb_structure_array(h, i, _ptr(a), C.byref(n))
            out[name] = a.reshape(-1, 8) if name == "tmeta" else a
        for i, name in ((21, "frag_tile"), (22, "hv_pt"), (23, "hv_ptr")):  # long tracks
            n = C.c_int64()
            L.gb_structure_array(h, i, None, C.byref(n))
            a = np.empty(n.value, dtype=np.int32)
            L.gb_structure_array(h, i, _ptr(a), C.byref(n))
            out[name] = a
        info = (C.c_int64 * 12)()
        L.gb_structure_info(h, info)
        out["info"] = dict(zip(INFO_KEYS, [int(v) for v in info]))
        nblk = int(n_cams) + int(n_pts)
        nnz = int(n_cams) + len(ci) + int(n_pts)
        cp = np.empty(nblk + 1, dtype=np.int64); ri = np.empty(nnz, dtype=np.int64); off = np.empty(nnz, dtype=np.int64)
        L.gb_structure_hessian(h, _ptr(cp), _ptr(ri), _ptr(off))
        out["hessian"] = (cp, ri, off)
        n = C.c_int64()
        L.gb_structure_schur(h, None, None, C.byref(n))
        scp = np.empty(int(n_cams) + 1, dtype=np.int64); sri = np.empty(n.value, dtype=np.int64)
        L.gb_structure_schur(h, _ptr(scp), _ptr(sri), C.byref(n))
        out["schur"] = (scp, sri)
        return out
    finally:
        L.gb_structure_destroy(h)

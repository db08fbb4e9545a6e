"""GPU (-m gpu): the GENERIC factor-graph path (include/graphite_b200_graph.h) through the C ABI.

  * the integer known-answer tests of the reference's tests/factor.cu (toy unary / binary factors), restated through the ABI;
  * a 6-dof pose graph (between factors 6/6/6 with Huber loss + precision matrices + activity levels, unary priors, fixed
    vertices) against the numpy oracle stage by stage and against runs of the unmodified reference (tests/golden/pose-graph*);
  * SE(2) pose graph with a user update callback; the BAL factor 2/9/3 through the generic path against the specialised
    BAL path and the reference's `pcg` golden run.
Factors are evaluated by USER kernels (tests/user_factor/graph_factors.cu) compiled against the public header only.
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, golden_json, golden_npz
from graphite_b200 import binding, graph as gg, synthetic
from oracle import oracle_graph as og

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


@pytest.fixture(scope="module")
def ctx(built):
    c = binding.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def ulib():
    """Compile the user's factor kernels the way a Graphite user would: nvcc + the public header."""
    src = os.path.join(ROOT, "tests", "user_factor", "graph_factors.cu")
    out = os.path.join(ROOT, "tests", "user_factor", "libgraph_factors.so")
    deps = [src, os.path.join(ROOT, "tests", "user_factor", "pose_residual.cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(out) < os.path.getmtime(d) for d in deps):
        subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a",
                               "--expt-relaxed-constexpr", "-shared", "-Xcompiler", "-fPIC", "-o", out, src])
    L = ctypes.CDLL(out)
    L.user_device_upload.restype = ctypes.c_void_p
    L.user_device_upload.argtypes = [ctypes.c_void_p, ctypes.c_longlong]
    L.user_device_free.argtypes = [ctypes.c_void_p]
    return L


def fn(lib, name):
    return ctypes.cast(getattr(lib, name), ctypes.c_void_p).value


class Dev:
    """Device copies of the user's per-factor data (observations), freed at the end of the test."""

    def __init__(self, lib):
        self.lib, self.ptrs, self.keep = lib, [], []

    def up(self, a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        p = self.lib.user_device_upload(a.ctypes.data_as(ctypes.c_void_p), a.nbytes)
        assert p
        self.ptrs.append(p)
        return p

    def close(self):
        for p in self.ptrs:
            self.lib.user_device_free(p)
        self.ptrs = []


class LinearUser(ctypes.Structure):
    _fields_ = [("E", ctypes.c_int), ("arity", ctypes.c_int), ("d", ctypes.c_int * 4), ("A", ctypes.c_void_p * 4), ("obs", ctypes.c_void_p)]


def linear_user(dev, A, obs):
    """A: list of E x d arrays (one per slot); the struct the user's linear kernel reads."""
    E = A[0].shape[0]
    u = LinearUser()
    u.E, u.arity = E, len(A)
    for s, a in enumerate(A):
        u.d[s] = a.shape[1]
        u.A[s] = dev.up(np.asarray(a, dtype=np.float64).T.reshape(-1))  # column-major
    u.obs = dev.up(np.asarray(obs, dtype=np.float64).reshape(-1))
    dev.keep.append(u)
    return ctypes.addressof(u)


# ------------------------------------------------------------------------------------------------------
# the reference's integer known-answer tests (tests/factor.cu), restated through the C ABI
# ------------------------------------------------------------------------------------------------------
def toy_graph(ctx, ulib, dev, factors, vertices, precision="f32-f32", fixed=None, scaling=False):
    """vertices: [(global id, (x, y))]; factors: list of (A list, [vertex ids per slot], obs, loss, delta, active)."""
    G = gg.Graph(ctx, precision)
    ids = [v[0] for v in vertices]
    vs = G.add_vertex_set(2, ids, fixed=fixed)
    suffix = "f32" if precision.startswith("f32") else "f64"
    for A, conn, obs, loss, delta, active in factors:
        idx = [[ids.index(g) for g in c] for c in conn]
        G.add_factor_set(A[0].shape[0], [vs] * len(A), idx, fn(ulib, "linear_factor_" + suffix), linear_user(dev, A, obs), active=active,
                         loss=loss, loss_delta=delta)
    G.set_vertices(vs, np.array([v[1] for v in vertices], dtype=np.float64))
    G.set_scaling(scaling)
    return G


UNARY = [np.array([[1.0, 0.0]])]            # tests/factor.cu:8-29
COUPLED = [np.array([[2.0, 3.0]])]          # :31-52
BINARY = [np.array([[1.0, 2.0]]), np.array([[3.0, 4.0]])]  # :54-82


def test_kat_compute_error_b_and_huber(ctx, ulib):
    """ComputeError (:141-158), ComputeB (:425-466), ComputeBHuberLoss (:468-509), Chi2HuberLoss (:758-784)."""
    dev = Dev(ulib)
    G = toy_graph(ctx, ulib, dev, [(UNARY, [[10], [10]], [2.5, 2.5], 0, 0.0, None)], [(10, (7.0, 0.0))])
    G.initialize(0)
    chi2 = G.linearize()
    assert np.array_equal(G.residuals(0).reshape(-1), np.float32([4.5, 4.5]))
    assert chi2 == 2 * 4.5 ** 2
    # one compute_b call of two identical factors: b0 = -2 * 4.5 (the reference test calls it twice on a pre-filled b)
    assert np.array_equal(G.gradient(), np.float32([-9.0, 0.0]))
    G.close()
    G = toy_graph(ctx, ulib, dev, [(UNARY, [[10], [10]], [2.5, 2.5], 1, 1.0, None)], [(10, (7.0, 0.0))])
    G.initialize(0)
    G.linearize()
    assert np.allclose(G.gradient(), [-2.0, 0.0], rtol=1e-6)  # each factor contributes -1 (dL = 1 / 4.5)
    G.close()
    G = toy_graph(ctx, ulib, dev, [(UNARY, [[10], [10]], [2.5, 6.5], 1, 1.0, None)], [(10, (7.0, 0.0))])
    G.initialize(0)
    assert G.linearize() == pytest.approx(8.25, rel=1e-6)  # rho(4.5^2) = 8, rho(0.5^2) = 0.25
    assert np.allclose(G.chi2_per_factor(0), [8.0, 0.25], rtol=1e-6)
    G.close()
    dev.close()


def test_kat_diagonals_scaling_and_products(ctx, ulib):
    """ComputeHessianBlockDiagonal (:511-555), ComputeHessianScalarDiagonal (:557-595), ScaleJacobiansAsync (:383-423),
    ComputeJvHuberLoss (:597-675), ComputeJtvHuberLoss (:677-756)."""
    dev = Dev(ulib)
    G = toy_graph(ctx, ulib, dev, [(COUPLED, [[10], [10]], [2.5, 2.5], 0, 0.0, None)], [(10, (7.0, 0.0))])
    G.initialize(0)
    G.linearize()
    assert np.array_equal(G.block_diagonal(0).reshape(-1), np.float32([8, 12, 12, 18]))  # 2 * J^T J, J = [2, 3]
    assert np.array_equal(G.scalar_diagonal(), np.float32([8, 18]))
    assert np.array_equal(G.jacobians(0, 0).reshape(-1), np.float32([2, 3, 2, 3]))
    G.close()
    # Jacobi scaling: s_j = 1 / (eps + sqrt(d_j)), J~ = J s (graph.hpp:254-281); the KAT's literal scales [2, 3] -> [4, 9]
    # are a direct kernel call in the reference; through the graph the scales are the computed ones
    G = toy_graph(ctx, ulib, dev, [(COUPLED, [[10], [10]], [2.5, 2.5], 0, 0.0, None)], [(10, (7.0, 0.0))], scaling=True)
    G.initialize(0)
    G.linearize()
    s = G.scales()
    assert np.allclose(s, 1.0 / np.sqrt([8.0, 18.0]), rtol=1e-6)
    assert np.allclose(G.jacobians(0, 0).reshape(2, 2), np.float32([2, 3]) * s, rtol=1e-6)
    assert np.allclose(G.scalar_diagonal(), [1.0, 1.0], rtol=1e-6)
    G.close()
    # J v and J^T dL P v with the Huber weight (delta = 1, residual 4.5: dL = 1 / 4.5)
    G = toy_graph(ctx, ulib, dev, [(UNARY, [[10], [10]], [2.5, 2.5], 1, 1.0, None)], [(10, (7.0, 0.0))])
    G.initialize(0)
    G.linearize()
    assert np.array_equal(G.jv(np.float32([3.0, 5.0])), np.float32([3.0, 3.0]))
    assert np.allclose(G.jtpv(np.float32([9.0, 9.0])), [4.0, 0.0], rtol=1e-6)
    G.close()
    dev.close()


def test_kat_hessian_structure_and_values(ctx, ulib):
    """ComputeHessian (:854-967): unary on v0 and v1 plus one binary (v0, v1): block coordinates (0,0) x2, (0,1), (1,1) x2,
    value offsets {0, 4, 8}, H = [5,8,8,13 | 3,6,4,8 | 13,18,18,25]."""
    dev = Dev(ulib)
    G = toy_graph(ctx, ulib, dev, [(COUPLED, [[10], [20]], [2.5, 3.5], 0, 0.0, None), (BINARY, [[10, 20]], [4.5], 0, 0.0, None)],
                  [(10, (7.0, 5.0)), (20, (11.0, 13.0))])
    info = G.initialize(0)
    assert (info["hessian_dim"], info["block_columns"], info["hessian_blocks"], info["hessian_values"]) == (4, 2, 3, 12)
    cp, ri, off = G.hessian_structure()
    assert cp.tolist() == [0, 1, 3] and ri.tolist() == [0, 0, 1] and off.tolist() == [0, 4, 8]
    G.linearize()
    assert np.array_equal(G.hessian_values(), np.float32([5, 8, 8, 13, 3, 6, 4, 8, 13, 18, 18, 25]))
    G.close()
    dev.close()


def test_fixed_vertices_and_inactive_factors(ctx, ulib):
    """set_fixed / set_active semantics (vertex.hpp:254-266, active.hpp:11-21, graph.hpp:171-210; tests/factor.cu:326-358,
    :640-675): fixed and unused vertices get no Hessian column, inactive factors contribute nothing, levels gate factors."""
    dev = Dev(ulib)
    verts = [(10, (7.0, 5.0)), (20, (11.0, 13.0)), (30, (1.0, 2.0))]
    facs = [(COUPLED, [[10], [20], [30]], [2.5, 3.5, 0.5], 0, 0.0, [0, 0, 1]), (BINARY, [[10, 20]], [4.5], 0, 0.0, None)]
    G = toy_graph(ctx, ulib, dev, facs, verts, precision="f64-f64", fixed=[0, 1, 0])
    info = G.initialize(0)
    # vertex 20 is fixed, vertex 30 is only referenced by a level-1 factor: one block column at level 0
    assert info["block_columns"] == 1 and info["hessian_dim"] == 2 and info["active_factors"] == 3
    assert G.vertex_columns(0).tolist() == [0, -1, -1]
    chi2 = G.linearize()
    r_un, r_bi = 2 * 7 + 3 * 5 - 2.5, 7 + 2 * 5 + 3 * 11 + 4 * 13 - 4.5
    r_20 = 2 * 11 + 3 * 13 - 3.5
    assert chi2 == pytest.approx(r_un ** 2 + r_20 ** 2 + r_bi ** 2, rel=1e-14)  # factors on fixed vertices still cost
    assert np.allclose(G.gradient(), [-(2 * r_un + 1 * r_bi), -(3 * r_un + 2 * r_bi)], rtol=1e-14)
    assert np.array_equal(G.jacobians(1, 1).reshape(-1), [0.0, 0.0])  # slot of the fixed vertex stays zero
    x, inf = G.solve(max_iterations=20, tolerance=1e-30)
    before = G.get_vertices(0).copy()
    traj, res = G.lm(iterations=5, initial_damping=1e-6, pcg_iterations=20, pcg_tolerance=1e-30)
    after = G.get_vertices(0)
    assert np.array_equal(after[1], before[1]) and np.array_equal(after[2], before[2])  # fixed / unused: untouched
    assert not np.array_equal(after[0], before[0]) and traj[-1, 1] < traj[0, 0]
    info1 = G.initialize(1)  # level 1: the third unary factor and with it vertex 30 become active
    assert info1["block_columns"] == 2 and info1["active_factors"] == 4 and G.vertex_columns(0).tolist() == [0, -1, 2]
    G.close()
    dev.close()


# ------------------------------------------------------------------------------------------------------
# 6-dof pose graph: oracle and reference parity
# ------------------------------------------------------------------------------------------------------
def pose_oracle(pg, level=0):
    O = og.GraphOracle()
    v = O.add_vertex_set(6, pg.ids, pg.poses, pg.fixed)
    O.add_factor_set(6, [v, v], pg.bt_idx, og.make_between6(pg.bt_meas), active=pg.bt_active, P=pg.bt_P, loss=1, delta=pg.huber)
    O.add_factor_set(6, [v], pg.pr_idx, og.make_prior6(pg.pr_meas))
    O.initialize(level)
    return O


def pose_graph_gpu(ctx, ulib, dev, pg, precision="f64-f64", level=0):
    G = gg.Graph(ctx, precision)
    sfx = "f32" if precision.startswith("f32") else "f64"
    v = G.add_vertex_set(6, pg.ids, fixed=pg.fixed)
    G.add_factor_set(6, [v, v], pg.bt_idx, fn(ulib, "between6_" + sfx), dev.up(pg.bt_meas), active=pg.bt_active, loss=1, loss_delta=pg.huber)
    G.add_factor_set(6, [v], pg.pr_idx.reshape(-1, 1), fn(ulib, "prior6_" + sfx), dev.up(pg.pr_meas))
    G.set_precision(0, pg.bt_P)
    G.set_vertices(v, pg.poses)
    G.initialize(level)
    return G


PROTO = dict(iterations=12, initial_damping=1e-4, pcg_iterations=30, pcg_tolerance=1e-10, rejection_ratio=5.0)
# the hard start needs a tighter linear solve to stay reproducible across implementations (a 1e-13 perturbation of the
# start moves this trajectory by 1e-10; with 30 PCG iterations and lambda = 1e-7 it is chaotic from the 4th step on)
PROTO_HARD = dict(iterations=14, initial_damping=1e-4, pcg_iterations=100, pcg_tolerance=1e-10, rejection_ratio=5.0)


@pytest.mark.parametrize("level", [0, 1])
def test_pose_graph_stages_match_oracle(ctx, ulib, level):
    pg = synthetic.pose_graph()
    dev = Dev(ulib)
    G = pose_graph_gpu(ctx, ulib, dev, pg, level=level)
    O = pose_oracle(pg, level)
    assert G.info["hessian_dim"] == O.dimH and G.info["block_columns"] == O.nblocks
    cp, ri, off = G.hessian_structure()
    assert np.array_equal(cp, O.colptr) and np.array_equal(ri, O.rowidx) and np.array_equal(off, O.offsets)
    assert np.array_equal(G.vertex_columns(0), O.V[0]["hoff"])
    chi2, ochi2 = G.linearize(), O.linearize()
    assert abs(chi2 - ochi2) / ochi2 < 1e-13
    assert rel(G.scales(), O.scales) < 1e-12 and rel(G.gradient(), O.b) < 1e-12
    assert rel(G.hessian_values(), O.hessian_values()) < 1e-12
    assert rel(G.loss_derivative(0), O.F[0]["dL"]) < 1e-12 and O.F[0]["dL"][O.F[0]["active"]].min() < 1.0  # Huber branch taken
    rng = np.random.default_rng(1)
    x, v = rng.normal(size=O.dimH), rng.normal(size=O.rows)
    jv = G.jv(x)
    mask = np.concatenate([np.repeat(f["active"], f["E"]) for f in O.F])
    assert rel(jv[mask], O.jv(x)[mask]) < 1e-12 and not jv[~mask].any()
    assert rel(G.jtpv(v), O.jtpv(v)) < 1e-12
    G.set_damping(1e-4)
    xg, inf = G.solve(max_iterations=30, tolerance=1e-10)
    xo, ko = O.solve(1e-4, False, 30, 1e-10, 5.0)
    assert inf["pcg_iterations"] == ko and rel(xg, xo) < 1e-9
    G.close()
    dev.close()


@pytest.mark.parametrize("case", ["pose-graph", "pose-graph-hard"])
def test_pose_graph_lm_matches_oracle(ctx, ulib, case):
    pg = synthetic.pose_graph() if case == "pose-graph" else synthetic.pose_graph_hard()
    proto = dict(PROTO) if case == "pose-graph" else dict(PROTO_HARD)
    dev = Dev(ulib)
    G = pose_graph_gpu(ctx, ulib, dev, pg)
    traj, res = G.lm(**proto)
    otab = pose_oracle(pg).lm(iterations=proto["iterations"], initial_damping=proto["initial_damping"],
                              pcg_iterations=proto["pcg_iterations"], pcg_tolerance=1e-10, rejection_ratio=5.0)
    n = min(len(traj), len(otab))
    assert n == proto["iterations"]
    assert np.array_equal(traj[:n, 0] == traj[:n, 1], otab[:n, 0] == otab[:n, 1])  # accept / reject decisions
    # far from the optimum (1.2 rad rotations, rejected steps) rounding differences are amplified ~1000x
    tol = 1e-9 if case == "pose-graph" else 1e-8
    assert np.abs(traj[:n, 1] - otab[:n, 1]).max() / otab[0, 0] < tol
    if case == "pose-graph":
        assert np.array_equal(traj[:n, 3], otab[:n, 3])
    if case == "pose-graph-hard":
        assert res["rejected"] > 0 and res["accepted"] > 0
    # runs are bit-reproducible (gathers in a fixed order, no atomics)
    G2 = pose_graph_gpu(ctx, ulib, dev, pg)
    traj2, _ = G2.lm(**proto)
    assert np.array_equal(traj, traj2) and np.array_equal(G.get_vertices(0), G2.get_vertices(0))
    G.close(); G2.close()
    dev.close()


def _golden(name):
    try:
        return golden_json(name + ".json"), golden_npz(name + ".npz")
    except FileNotFoundError:
        pytest.skip(f"golden {name} not generated yet (oracle/make_golden_pose.py on the GPU box)")


@pytest.mark.parametrize("level", [0, 1])
def test_pose_graph_matches_reference(ctx, ulib, level):
    """Against the unmodified reference (oracle/_ref/ref_pose on a B200): structure bit-exact, first linearisation 1e-12,
    LM trajectory 1e-9 per iteration, final cost 1e-6 (north_star tolerances)."""
    js, z = _golden(f"pose-graph__pcg__FP64-FP64__level{level}")
    pg = synthetic.pose_graph()
    dev = Dev(ulib)
    G = pose_graph_gpu(ctx, ulib, dev, pg, level=level)
    cp, ri, off = G.hessian_structure()
    assert np.array_equal(cp, z["H_colptr"]) and np.array_equal(ri, z["H_rowidx"]) and np.array_equal(off, z["H_offsets"])
    assert np.array_equal(G.vertex_columns(0), z["columns"])
    chi2 = G.linearize()
    assert abs(chi2 - js["initial_chi2_17g"]) / chi2 < 1e-12
    assert rel(G.scales(), z["scales"]) < 1e-12 and rel(G.gradient(), z["b"]) < 1e-11
    assert rel(G.hessian_values(), z["H_values"]) < 1e-11
    traj, res = G.lm(**PROTO)
    ref = np.array(js["table"])
    assert np.array_equal(traj[:, 0] == traj[:, 1], ref[:, 1] == ref[:, 2])
    assert np.abs(traj[:, 1] - ref[:, 2]).max() / ref[0, 1] < 1e-9
    assert abs(traj[-1, 1] - js["final_chi2"]) / js["final_chi2"] < 1e-6
    assert rel(G.get_vertices(0).reshape(-1), z["final_poses"]) < 1e-6
    G.close()
    dev.close()


def test_pose_graph_hard_matches_reference(ctx, ulib):
    js, _ = _golden("pose-graph-hard__pcg__FP64-FP64__level0")
    dev = Dev(ulib)
    G = pose_graph_gpu(ctx, ulib, dev, synthetic.pose_graph_hard())
    traj, res = G.lm(**PROTO_HARD)
    ref = np.array(js["table"])
    assert np.array_equal(traj[:, 0] == traj[:, 1], ref[:, 1] == ref[:, 2]) and res["rejected"] > 0
    assert np.abs(traj[:, 1] - ref[:, 2]).max() / ref[0, 1] < 1e-8  # see test_pose_graph_lm_matches_oracle
    assert abs(traj[-1, 1] - js["final_chi2"]) / js["final_chi2"] < 1e-6
    G.close()
    dev.close()


@pytest.mark.parametrize("precision,tag", [("f32-f32", "FP32-FP32"), ("f64-f32", "FP64-FP32")])
def test_pose_graph_low_precision_matches_reference(ctx, ulib, precision, tag):
    """FP32 / mixed storage: 1e-4 against the reference's own run of the same precision (north_star)."""
    js, z = _golden(f"pose-graph__pcg__{tag}__level0")
    dev = Dev(ulib)
    G = pose_graph_gpu(ctx, ulib, dev, synthetic.pose_graph(), precision=precision)
    chi2 = G.linearize()
    assert abs(chi2 - js["initial_chi2_17g"]) / chi2 < 1e-5
    assert rel(G.scales(), z["scales"]) < 1e-4 and rel(G.gradient(), z["b"]) < 1e-4
    traj, _ = G.lm(**PROTO)
    ref = np.array(js["table"])
    assert np.abs(traj[:, 1] - ref[:, 2]).max() / ref[0, 1] < 1e-4
    assert abs(traj[-1, 1] - js["final_chi2"]) / js["final_chi2"] < 1e-4
    G.close()
    dev.close()


def test_pose_graph_bf16_storage(ctx, ulib):
    """bf16 Jacobian storage with FP64 accumulation (types.hpp:10-19): converges to the FP64 optimum within bf16's resolution."""
    pg = synthetic.pose_graph()
    dev = Dev(ulib)
    G = pose_graph_gpu(ctx, ulib, dev, pg, precision="f64-bf16")
    G64 = pose_graph_gpu(ctx, ulib, dev, pg)
    traj, _ = G.lm(**PROTO)
    t64, _ = G64.lm(**PROTO)
    assert traj[0, 0] == pytest.approx(t64[0, 0], rel=1e-12)  # residuals are T precision
    assert abs(traj[-1, 1] - t64[-1, 1]) / t64[-1, 1] < 2e-2 and traj[-1, 1] < 0.3 * traj[0, 0]
    G.close(); G64.close()
    dev.close()


# ------------------------------------------------------------------------------------------------------
# SE(2) with a user update callback; BAL 2/9/3 through the generic path
# ------------------------------------------------------------------------------------------------------
def test_se2_pose_graph_with_update_callback(ctx, ulib):
    rng = np.random.Generator(np.random.PCG64(5))
    n = 30
    k = np.arange(n)
    gt = np.stack([3 * np.cos(0.4 * k), 3 * np.sin(0.4 * k), 0.4 * k + np.pi / 2], axis=1)
    gt[:, 2] = [og.wrap_pi(a) for a in gt[:, 2]]
    edges = [(i, i + 1) for i in range(n - 1)] + [(i, i + 7) for i in range(0, n - 7, 2)]
    meas = np.zeros((len(edges), 3))
    for e, (i, j) in enumerate(edges):
        c, s = np.cos(gt[i, 2]), np.sin(gt[i, 2])
        d = gt[j, :2] - gt[i, :2]
        meas[e] = [c * d[0] + s * d[1], -s * d[0] + c * d[1], og.wrap_pi(gt[j, 2] - gt[i, 2])]
    meas += rng.normal(0, 0.01, meas.shape)
    init = gt + rng.normal(0, 0.1, gt.shape)
    init[0] = gt[0]
    fixed = np.zeros(n, np.uint8); fixed[0] = 1
    ids = np.arange(n) * 2 + 5

    def upd(x, d):
        return np.array([x[0] + d[0], x[1] + d[1], og.wrap_pi(x[2] + d[2])])

    O = og.GraphOracle()
    v = O.add_vertex_set(3, ids, init, fixed, update=upd)
    O.add_factor_set(3, [v, v], edges, og.make_se2(meas))
    O.initialize(0)
    otab = O.lm(iterations=10, initial_damping=1e-4, pcg_iterations=40, pcg_tolerance=1e-12)
    dev = Dev(ulib)
    G = gg.Graph(ctx, "f64-f64")
    gv = G.add_vertex_set(3, ids, fixed=fixed)
    G.add_factor_set(3, [gv, gv], edges, fn(ulib, "se2_between_f64"), dev.up(meas))
    G.set_update(gv, fn(ulib, "se2_update_f64"))
    G.set_vertices(gv, init)
    G.initialize(0)
    traj, res = G.lm(iterations=10, initial_damping=1e-4, pcg_iterations=40, pcg_tolerance=1e-12)
    assert np.abs(traj[:, 1] - otab[:, 1]).max() / otab[0, 0] < 1e-9
    assert np.array_equal(traj[:5, 3], otab[:5, 3])  # (at the converged optimum rz sits at the tolerance: counts are rounding)
    assert rel(G.get_vertices(gv), O.V[0]["x"]) < 1e-9
    assert np.abs(G.get_vertices(gv)[:, 2]).max() <= np.pi  # the callback wrapped the angles
    assert traj[-1, 1] < 1e-2 * traj[0, 0]
    G.close()
    dev.close()


def test_bal_factor_through_the_generic_path(ctx, ulib):
    """The BAL reprojection factor as a generic 2/9/3 factor set (cameras first, points eliminated-ordered): same block
    order, gradient, scales and Hessian as the specialised BAL path, and the LM trajectory of the reference's `pcg` run."""
    prob = synthetic.make_named("ladybug-49")
    dev = Dev(ulib)
    G = gg.Graph(ctx, "f64-f64")
    vc = G.add_vertex_set(9, np.arange(prob.n_cams))
    vp = G.add_vertex_set(3, prob.n_cams + np.arange(prob.n_pts), eliminate=True)
    G.add_factor_set(2, [vc, vp], np.stack([prob.cam_idx, prob.pt_idx], axis=1), fn(ulib, "bal_graph_factor_f64"), dev.up(prob.obs))
    G.set_vertices(vc, prob.cams)
    G.set_vertices(vp, prob.pts)
    info = G.initialize(0)
    P = binding.problem_from_bal(ctx, prob, "f64-f64")
    cp, ri, off = P.hessian_structure()
    gcp, gri, goff = G.hessian_structure()
    assert np.array_equal(cp, gcp) and np.array_equal(ri, gri) and np.array_equal(off, goff)
    chi2, gchi2 = P.linearize(), G.linearize()
    assert abs(chi2 - gchi2) / chi2 < 1e-14
    assert rel(G.scales(), P.scales()) < 1e-13 and rel(G.gradient(), P.gradient()) < 1e-12
    assert rel(G.hessian_values(), P.hessian_values()) < 1e-12
    js = golden_json("ladybug-49__pcg__FP64-FP64.json")
    ref = np.array(js["table"])
    traj, res = G.lm(iterations=len(ref), initial_damping=1e-4, pcg_iterations=10, pcg_tolerance=1.0, rejection_ratio=5.0)
    assert np.array_equal(traj[:, 0] == traj[:, 1], ref[:, 1] == ref[:, 2])
    assert np.abs(traj[:, 1] - ref[:, 2]).max() / ref[0, 1] < 1e-9
    assert abs(traj[-1, 1] - js["final_chi2"]) / js["final_chi2"] < 1e-6
    P.close(); G.close()
    dev.close()


def test_graph_errors_are_reported(ctx, ulib):
    dev = Dev(ulib)
    G = gg.Graph(ctx, "f64-f64")
    v = G.add_vertex_set(2, [1, 2])
    with pytest.raises(binding.GraphiteB200Error, match="out of range"):
        G.add_factor_set(1, [v], [[5]], fn(ulib, "linear_factor_f64"), linear_user(dev, UNARY, [0.0]))
    with pytest.raises(binding.GraphiteB200Error, match="callback missing"):
        G.add_factor_set(1, [v], [[0]], None)
    G.add_factor_set(1, [v], [[0], [1]], fn(ulib, "failing_graph_factor"))
    with pytest.raises(binding.GraphiteB200Error, match="not initialised"):
        G.linearize()
    G.initialize(0)
    with pytest.raises(binding.GraphiteB200Error, match="vertex values missing"):
        G.linearize()
    G.set_vertices(v, np.zeros((2, 2)))
    with pytest.raises(binding.GraphiteB200Error, match="returned 7"):
        G.linearize()
    with pytest.raises(binding.GraphiteB200Error):
        gg.Graph(ctx, "f32-f64")
    G2 = gg.Graph(ctx, "f64-f64")
    G2.add_vertex_set(2, [1, 1])
    G2.add_factor_set(1, [0], [[0]], fn(ulib, "linear_factor_f64"), linear_user(dev, UNARY, [0.0]))
    with pytest.raises(binding.GraphiteB200Error, match="duplicate global vertex id"):
        G2.initialize(0)
    G.close(); G2.close()
    dev.close()

"""compute-sanitizer target: a few LM iterations of small problems through every kernel of the library - the BAL path (FP64
two workers, FP32 three workers, both solvers, explicit and direct Schur, Huber + precision matrices, fixed vertices, device
structure build, long tracks cut into fragment tiles, the Jacobian-free and the Jacobian-streaming back-substitution) and
the generic factor-graph path (pose graph with the user kernels of tests/user_factor)."""
import ctypes
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from graphite_b200 import binding, graph as gg, synthetic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
case = sys.argv[1] if len(sys.argv) > 1 else "ladybug-49"
prob = synthetic.make_named(case)
ctx = binding.Context(0)
P = binding.problem_from_bal(ctx, prob, "f64-f64")
t, r = P.lm(iterations=3)
P.set_loss("huber", 20.0); P.set_precision(synthetic.precision_matrices(prob.n_obs))
P.set_vertices(prob.cams, prob.pts)
t2, _ = P.lm(iterations=2)
P.set_vertices(prob.cams, prob.pts)
t3, _ = P.lm(iterations=2, solver="pcg")
P.linearize(); P.set_damping(1e-3); v = P.schur_values()
P.set_loss("default"); P.set_precision(None)
fc = np.zeros(prob.n_cams, np.uint8); fc[:2] = 1
fp = np.zeros(prob.n_pts, np.uint8); fp[::50] = 1
P.set_fixed(fc, fp)
P.set_vertices(prob.cams, prob.pts)
t4, _ = P.lm(iterations=2)
t5, _ = P.lm(iterations=2, schur_mode="explicit")
t6, _ = P.lm(iterations=1, solver="direct-schur")
P.close()
P32 = binding.problem_from_bal(ctx, prob, "f32-f32")
t7, _ = P32.lm(iterations=3)
P32.close()
print("bal ok", t[-1, 1], t2[-1, 1], t3[-1, 1], v.shape, t4[-1, 1], t5[-1, 1], t6[-1, 1], t7[-1, 1])

# long tracks: fragment tiles, phase H of k_pcg_solve, k_frag_sum / k_frag_dots; a 20-iteration solve takes the
# Jacobian-streaming back-substitution, the 10-iteration ones the kept point sums
lt = synthetic.make_named("long-tracks")
for kw in ({}, dict(tile_size=24, slot_cap=40)):
    PL = binding.problem_from_bal(ctx, lt, "f64-f64", **kw)
    l1, _ = PL.lm(iterations=2)
    l2, _ = PL.lm(iterations=1, pcg_iterations=20)
    l3, _ = PL.lm(iterations=1, solver="pcg")
    l4, _ = PL.lm(iterations=1, schur_mode="explicit")
    PL.close()
    print("long tracks ok", kw, l1[-1, 1], l2[-1, 1], l3[-1, 1], l4[-1, 1])
PL = binding.problem_from_bal(ctx, lt, "f32-f32")
l5, _ = PL.lm(iterations=2)
PL.close()
print("long tracks f32 ok", l5[-1, 1])

# generic path
src = os.path.join(ROOT, "tests", "user_factor", "graph_factors.cu")
lib = os.path.join(ROOT, "tests", "user_factor", "libgraph_factors.so")
if not os.path.exists(lib):
    subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a",
                           "--expt-relaxed-constexpr", "-shared", "-Xcompiler", "-fPIC", "-o", lib, src])
U = ctypes.CDLL(lib)
U.user_device_upload.restype = ctypes.c_void_p
U.user_device_upload.argtypes = [ctypes.c_void_p, ctypes.c_longlong]
fn = lambda name: ctypes.cast(getattr(U, name), ctypes.c_void_p).value
up = lambda a: U.user_device_upload(np.ascontiguousarray(a, dtype=np.float64).ctypes.data_as(ctypes.c_void_p), a.size * 8)
pg = synthetic.pose_graph()
for prec, sfx in (("f64-f64", "f64"), ("f32-f32", "f32"), ("f64-f32", "f64")):
    G = gg.Graph(ctx, prec)
    vs = G.add_vertex_set(6, pg.ids, fixed=pg.fixed)
    G.add_factor_set(6, [vs, vs], pg.bt_idx, fn("between6_" + sfx), up(pg.bt_meas), active=pg.bt_active, loss=1, loss_delta=pg.huber)
    G.add_factor_set(6, [vs], pg.pr_idx.reshape(-1, 1), fn("prior6_" + sfx), up(pg.pr_meas))
    G.set_precision(0, pg.bt_P)
    G.set_vertices(vs, pg.poses)
    G.initialize(0)
    tg, _ = G.lm(iterations=3, pcg_iterations=30, pcg_tolerance=1e-10)
    G.linearize(); h = G.hessian_values()
    print("graph ok", prec, tg[-1, 1], h.shape)
    G.close()

"""Repeat the LM run N times on one problem and check bit-identical trajectories (race detector for the TMA pipelines)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graphite_b200 import binding, synthetic

name = sys.argv[1] if len(sys.argv) > 1 else "venice-1778"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
prec = sys.argv[3] if len(sys.argv) > 3 else "f64-f64"
prob = synthetic.make_named(name)
ctx = binding.Context(0)
P = binding.problem_from_bal(ctx, prob, prec)
ref = None
bad = 0
for i in range(reps):
    P.set_vertices(prob.cams, prob.pts)
    traj, res = P.lm(iterations=30)
    c, p = P.get_vertices()
    if ref is None:
        ref = (traj.copy(), c.copy(), p.copy())
    else:
        same = np.array_equal(ref[0], traj) and np.array_equal(ref[1], c) and np.array_equal(ref[2], p)
        if not same:
            bad += 1
            d = np.abs(ref[0][:, 1] - traj[:, 1]) / ref[0][:, 1]
            print(f"run {i}: DIFFERS, first differing iteration {int(np.argmax(d > 0))}, max rel {d.max():.3e}")
print(f"{name} {prec}: {reps} runs, {bad} differing; final chi2 {ref[0][-1, 1]:.12g}")
sys.exit(1 if bad else 0)

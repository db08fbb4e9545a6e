"""compute-sanitizer target: a few LM iterations of a small problem through every kernel of the path (both solvers)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from graphite_b200 import binding, synthetic
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
print("ok", t[-1, 1], t2[-1, 1], t3[-1, 1], v.shape)

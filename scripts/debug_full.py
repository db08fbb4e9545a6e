import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from graphite_b200 import binding, synthetic
from oracle.binding import Oracle, default_options
ctx = binding.Context(0)
for case in ["schur-fixture", "ladybug-49"]:
    prob = synthetic.schur_fixture() if case == "schur-fixture" else synthetic.make_named(case)
    P = binding.problem_from_bal(ctx, prob, "f64-f64")
    O = Oracle(prob)
    P.linearize(); O.linearize()
    for mu, ident in [(1e-4, False), (2.0, False), (1e-2, True)]:
        for iters in [1, 2, 3, 5, 10, 25]:
            P.set_damping(mu, ident)
            d, info = P.solve(iters, 1e-30, 1e30, solver="pcg")
            od, ok = O.solve(mu, default_options(solver=2, pcg_iterations=iters, pcg_tolerance=1e-30, rejection_ratio=1e30, use_identity=int(ident)))
            nc9 = 9 * prob.n_cams
            rc = np.abs(d[:nc9] - od[:nc9]).max() / np.abs(od[:nc9]).max()
            rp = np.abs(d[nc9:] - od[nc9:]).max() / np.abs(od[nc9:]).max()
            print(case, mu, ident, iters, info["pcg_iterations"], ok, "rel cam %.2e pt %.2e" % (rc, rp), flush=True)

"""Ill-conditioned robust runs (Huber loss / precision matrices, lambda down to 1e-11): deviation of the CUDA path from
the reference's golden run per iteration, next to the run's own sensitivity: the same run from vertices perturbed by
1e-12 relative.  No implementation can be expected to agree with another better than that."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from graphite_b200 import binding, synthetic
G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
ctx = binding.Context(0)
for name, solver, tag, huber, w in [("trafalgar-257", "pcg-schur", "trafalgar-257__pcg-schur__FP64-FP64__huber20__weights", 20.0, True),
                            ("ladybug-49", "pcg-schur", "ladybug-49__pcg-schur__FP64-FP64__huber20__weights", 20.0, True),
                            ("ladybug-49", "pcg-schur", "ladybug-49__pcg-schur__FP64-FP64__huber20", 20.0, False),
                            ("ladybug-49", "pcg-schur", "ladybug-49__pcg-schur__FP64-FP64__weights", 0.0, True),
                            ("ladybug-49", "pcg", "ladybug-49__pcg__FP64-FP64__huber20__weights", 20.0, True)]:
    t = np.array(json.load(open(os.path.join(G, tag + ".json")))["table"])
    prob = synthetic.make_named(name)
    P = binding.problem_from_bal(ctx, prob, "f64-f64")
    if huber > 0: P.set_loss("huber", huber)
    if w: P.set_precision(synthetic.precision_matrices(prob.n_obs))
    traj, res = P.lm(iterations=len(t), solver=solver)
    n = min(len(t), len(traj))
    r = np.abs(traj[:n, 1] - t[:n, 2]) / t[:n, 2]
    print(tag)
    print("  vs ref ", " ".join(f"{v:.1e}" for v in r))
    for eps in (1e-13, 1e-12):
        rng = np.random.default_rng(0)
        P.set_vertices(prob.cams * (1 + eps * rng.standard_normal(prob.cams.shape)), prob.pts * (1 + eps * rng.standard_normal(prob.pts.shape)))
        tp, _ = P.lm(iterations=len(t), solver=solver)
        m = min(n, len(tp))
        s = np.abs(tp[:m, 1] - traj[:m, 1]) / traj[:m, 1]
        print(f"  sens {eps:g}", " ".join(f"{v:.1e}" for v in s))
    P.close()
ctx.close()

"""FP32 / mixed diagnostics: deviation of the CUDA path from the reference's golden run per iteration, next to the
reference's own run-to-run spread when a *.run2.json exists (the reference sums with float atomics)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from graphite_b200 import binding, synthetic

G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
ctx = binding.Context(0)
for case in sys.argv[1:] or ["ladybug-49", "trafalgar-257", "dubrovnik-356", "venice-1778"]:
    g = json.load(open(os.path.join(G, f"{case}__pcg-schur__FP32-FP32.json")))
    t = np.array(g["table"])
    prob = synthetic.make_named(case)
    P = binding.problem_from_bal(ctx, prob, "f32-f32")
    traj, res = P.lm(iterations=len(t))
    n = min(len(traj), len(t))
    r = np.abs(traj[:n, 1] - t[:n, 2]) / t[:n, 2]
    line = {"case": case, "n": n, "ours_final": float(traj[-1, 1]), "ours_best": float(traj[:, 1].min()), "ref_final": g["final_chi2"],
            "rel_final": abs(traj[-1, 1] - g["final_chi2"]) / g["final_chi2"], "rel_per_iteration": [float(f"{v:.2e}") for v in r],
            "lambda_ref": [float(f"{v:.2e}") for v in t[:n, 3]], "lambda_ours": [float(f"{v:.2e}") for v in traj[:n, 2]]}
    p2 = os.path.join(G, f"{case}__pcg-schur__FP32-FP32.run2.json")
    if os.path.exists(p2):
        t2 = np.array(json.load(open(p2))["table"])
        m = min(len(t2), len(t))
        line["ref_spread"] = [float(f"{v:.2e}") for v in np.abs(t2[:m, 2] - t[:m, 2]) / t[:m, 2]]
    print(json.dumps(line))
    P.close()
ctx.close()

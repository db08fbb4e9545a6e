"""Explicit or implicit Schur complement?  Measured per problem shape: time of one gb_solve (prepare + [build S] + PCG with
EVERY iteration executed: tolerance 0, huge rejection ratio) for k = 5 .. 80 PCG iterations in both modes.
A linear fit time = a + b k per mode gives the set-up cost a (explicit: the S build) and the per-iteration cost b; the
crossover iteration count k* = (a_e - a_i) / (b_i - b_e) is what the library's GB_SCHUR_AUTO rule encodes."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graphite_b200 import binding, synthetic  # noqa: E402

ctx = binding.Context(0)
ks = [5, 10, 20, 40, 80]
for case in sys.argv[1:] or ["ladybug-49", "trafalgar-257", "dubrovnik-356", "venice-1778"]:
    prob = synthetic.make_named(case)
    P = binding.problem_from_bal(ctx, prob, "f64-f64")
    P.linearize()
    P.set_damping(1e-2)
    out = {"case": case, "shape": prob.shape()}
    for mode in ("implicit", "explicit"):
        ts = []
        for k in ks:
            best = 1e9
            for rep in range(4):
                t0 = time.perf_counter()
                d, info = P.solve(k, 0.0, 1e300, want_delta=False, schur_mode=mode)
                best = min(best, time.perf_counter() - t0)
            assert info["pcg_iterations"] == k, info
            ts.append(best * 1e3)
        b, a = np.polyfit(ks, ts, 1)
        out[mode] = {"ms": [round(v, 4) for v in ts], "setup_ms": round(float(a), 4), "ms_per_iteration": round(float(b), 5)}
    cp, ri = P.schur_structure()
    out["schur_blocks"] = int(len(ri))
    di, de = out["implicit"], out["explicit"]
    if di["ms_per_iteration"] > de["ms_per_iteration"]:
        out["crossover_iterations"] = round((de["setup_ms"] - di["setup_ms"]) / (di["ms_per_iteration"] - de["ms_per_iteration"]), 1)
    else:
        out["crossover_iterations"] = None
    print(json.dumps(out), flush=True)
    P.close()
ctx.close()

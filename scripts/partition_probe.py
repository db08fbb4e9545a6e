"""Single GPU: time every stage on rank 0's share of an N-way point partition (no communication), N = 1, 2, 4, 8,
for several super-tile sizes (waves of 148 CTAs).  Shows how far the kernels themselves scale when the per-GPU
problem shrinks; the collectives come on top."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graphite_b200 import binding, synthetic  # noqa: E402
from graphite_b200.distributed import partition_by_point  # noqa: E402

case = sys.argv[1] if len(sys.argv) > 1 else "venice-1778"
waves = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0]
prob = synthetic.make_named(case)
ctx = binding.Context(0)
names = ["lin", "prep", "prod+red", "backsub", "cost", "prod", "red", "preptiles"]
for n in [int(v) for v in (sys.argv[3].split(",") if len(sys.argv) > 3 else "1,2,4,8".split(","))]:
    part = partition_by_point(prob, n, 0)
    for w in waves:
        st_obs = 0 if w == 0 else (part.n_obs + 148 * w - 1) // (148 * w)
        P = binding.problem_from_bal(ctx, part, "f64-f64", partition=n > 1, super_tile_observations=st_obs)
        info = P.info()
        ms = [P.time_stage(s, 20) for s in range(8)]
        print(f"N={n} waves={w}: tiles {info['n_tiles']} super-tiles {info['n_super_tiles']} rows {info['n_partial_rows']} | "
              + " | ".join(f"{nm} {v * 1e3:.1f}" for nm, v in zip(names, ms)), flush=True)
        P.close()

"""ncu target: one pass of every stage on the bench workload (used with -k regex:<kernel> -c 1)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graphite_b200 import binding, synthetic  # noqa: E402

case = sys.argv[1] if len(sys.argv) > 1 else "venice-1778"
prec = sys.argv[2] if len(sys.argv) > 2 else "f64-f64"
prob = synthetic.make_named(case)
ctx = binding.Context(0)
P = binding.problem_from_bal(ctx, prob, prec)
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
traj, res = P.lm(iterations=iters)
print(traj)

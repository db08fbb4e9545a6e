"""Single GPU: a few LM iterations on rank 0's share of an N-way point partition as a stand-alone problem (no exchange).
Run under the ncu launch list: per-kernel device times at the per-GPU size of an N-GPU run."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graphite_b200 import binding, synthetic  # noqa: E402
from graphite_b200.distributed import partition_by_point  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
case = sys.argv[2] if len(sys.argv) > 2 else "venice-1778"
prob = synthetic.make_named(case)
part = partition_by_point(prob, n, 0)
ctx = binding.Context(0)
P = binding.problem_from_bal(ctx, part, "f64-f64", partition=n > 1)
traj, res = P.lm(iterations=10)
print({k: round(1e3 * v / 10, 4) for k, v in res.items() if k.startswith("seconds_")})

"""Does the staged observation upload overlap the LM iteration?  Times lm alone, the upload alone, and both."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from graphite_b200 import binding, synthetic
prob = synthetic.make_named("venice-1778")
ctx = binding.Context(0)
P = binding.problem_from_bal(ctx, prob, "f64-f64")
h_obs = torch.from_numpy(np.ascontiguousarray(prob.obs)).pin_memory()
h_c = torch.from_numpy(np.ascontiguousarray(prob.cams)).pin_memory(); h_p = torch.from_numpy(np.ascontiguousarray(prob.pts)).pin_memory()
P.lm(iterations=3)
def t(fn, n=10):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
def lm(): P.lm(iterations=1, initial_damping=1.0, resume=True)
def up_sync(): P.set_observations_raw(h_obs.data_ptr())
k = [0]
def stage_only():
    P.stage_observations_async(h_obs.data_ptr(), k[0] % 2); k[0] += 1
def both():
    P.stage_observations_async(h_obs.data_ptr(), k[0] % 2); k[0] += 1
    P.lm(iterations=1, initial_damping=1.0, resume=True)
print("lm alone          %.3f ms" % t(lm))
print("sync upload alone %.3f ms" % t(up_sync)); P.set_vertices(prob.cams, prob.pts); P.lm(iterations=1)
print("stage alone       %.3f ms (returns at once; the sync at the end of the loop waits for the copies)" % t(stage_only))
print("stage + lm        %.3f ms" % t(both))

import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from graphite_b200 import binding, synthetic
prob = synthetic.make_named("venice-1778")
ctx = binding.Context(0)
P = binding.problem_from_bal(ctx, prob, "f64-f64")
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
h_obs, h_c, h_p = pin(prob.obs), pin(prob.cams), pin(prob.pts)
oc, op = torch.empty_like(h_c).pin_memory(), torch.empty_like(h_p).pin_memory()
tw, rw = P.lm(iterations=3)
P.get_vertices_raw(oc.data_ptr(), op.data_ptr())
mu, nu = rw["final_damping"], rw["final_nu"]
P.stage_observations_async(h_obs.data_ptr(), 0); P.stage_observations_async(h_obs.data_ptr(), 1)
torch.cuda.synchronize()
def run(variant):
    global mu, nu
    for k in range(3):
        P.stage_observations_async(h_obs.data_ptr(), k % 2); torch.cuda.synchronize()   # something to commit
        ts = [time.perf_counter()]
        P.commit_observations(k % 2)
        ts.append(time.perf_counter())
        if variant in ("base", "verts_resume"): P.set_vertices_raw(oc.data_ptr(), op.data_ptr())
        if variant == "verts_resume": P.linearize()
        ts.append(time.perf_counter())
        P.stage_observations_async(h_obs.data_ptr(), (k + 1) % 2)
        ts.append(time.perf_counter())
        resume = variant in ("resume_noverts", "verts_resume")
        if variant == "resume_noverts" and k == 0: P.linearize(); P.stage_observations_async(h_obs.data_ptr(), (k + 1) % 2)
        tj, rj = P.lm(iterations=1, initial_damping=mu, initial_nu=nu, resume=resume)
        ts.append(time.perf_counter())
        P.get_vertices_raw(oc.data_ptr(), op.data_ptr()); ts.append(time.perf_counter())
        d = np.diff(ts) * 1e3
        print(variant, k, "commit %.2f setv(+lin) %.2f stage %.2f lm %.2f (device %.2f) getv %.2f" % (d[0], d[1], d[2], d[3], rj["seconds_total"] * 1e3, d[4]))
        torch.cuda.synchronize()
for v in ["base", "resume_noverts", "noverts_lin", "verts_resume"]:
    run(v)

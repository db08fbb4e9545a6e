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
def bench_like(n=20, defer=True):
    """The overlapped e2e loop of bench.py, with per-phase host wall times."""
    global mu, nu
    P.set_vertices_raw(h_c.data_ptr(), h_p.data_ptr())
    tw, rw = P.lm(iterations=3)
    P.get_vertices_raw(oc.data_ptr(), op.data_ptr())
    mu, nu = rw["final_damping"], rw["final_nu"]
    P.stage_observations_async(h_obs.data_ptr(), 0)
    torch.cuda.synchronize()
    acc = np.zeros(5); dev = 0.0
    t00 = time.perf_counter()
    for k in range(n):
        ts = [time.perf_counter()]
        P.commit_observations(k % 2); ts.append(time.perf_counter())
        P.set_vertices_raw(oc.data_ptr(), op.data_ptr()); ts.append(time.perf_counter())
        P.stage_observations_async(h_obs.data_ptr(), (k + 1) % 2); ts.append(time.perf_counter())
        tj, rj = P.lm(iterations=1, initial_damping=mu, initial_nu=nu, defer_final_linearize=defer); ts.append(time.perf_counter())
        P.get_vertices_raw(oc.data_ptr(), op.data_ptr()); ts.append(time.perf_counter())
        mu, nu = rj["final_damping"], rj["final_nu"]
        acc += np.diff(ts) * 1e3; dev += rj["seconds_total"] * 1e3
    torch.cuda.synchronize()
    tot = (time.perf_counter() - t00) * 1e3 / n
    print("bench-like defer=%s: per step %.2f ms | commit %.2f setv %.2f stage %.2f lm %.2f (device %.2f) getv %.2f" % ((defer, tot) + tuple(acc / n)[:4] + (dev / n, acc[4] / n)))
for v in ["base"]:
    run(v)
bench_like(20, True)
bench_like(20, False)


import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from graphite_b200 import binding, synthetic
from oracle.binding import Oracle, default_options
ctx = binding.Context(0)
prob = synthetic.make_named("trafalgar-257")
Pm = synthetic.precision_matrices(prob.n_obs)
O = Oracle(prob); O.set_robust("huber", 20.0, Pm)
O.lm(default_options(iterations=1))
c1, p1 = O.params()
prob2 = synthetic.BALProblem(prob.cam_idx, prob.pt_idx, prob.obs, c1.reshape(-1, 9), p1.reshape(-1, 3), "after-step-0")
P = binding.problem_from_bal(ctx, prob2, "f64-f64"); P.set_loss("huber", 20.0); P.set_precision(Pm)
O2 = Oracle(prob2); O2.set_robust("huber", 20.0, Pm)
chi2 = P.linearize(); ochi2, osc, ob = O2.linearize()
print("chi2", chi2, ochi2, "b rel", np.abs(P.gradient()-ob).max()/np.abs(ob).max())
mu = 1e-4/3
P.set_damping(mu)
d, info = P.solve(); od, ok = O2.solve(mu)
print("delta rel", np.abs(d-od).max()/np.abs(od).max(), info, ok)
new_chi2, _ = P.try_step()
# oracle: apply its own delta and evaluate
sc = osc
cams2 = c1.reshape(-1) + od[:9*prob.n_cams]*sc[:9*prob.n_cams]
pts2 = p1.reshape(-1) + od[9*prob.n_cams:]*sc[9*prob.n_cams:]
O3 = Oracle(synthetic.BALProblem(prob.cam_idx, prob.pt_idx, prob.obs, cams2.reshape(-1,9), pts2.reshape(-1,3), "x")); O3.set_robust("huber", 20.0, Pm)
r, c3 = O3.residuals()
print("cost after step: gpu %.10f oracle %.10f rel %.2e" % (new_chi2, c3, abs(new_chi2-c3)/c3))
# sensitivity: perturb the parameters by 1e-13 relative and redo the oracle step
rng = np.random.default_rng(0)
c1p = c1.reshape(-1)*(1+1e-13*rng.normal(size=c1.size)); p1p = p1.reshape(-1)*(1+1e-13*rng.normal(size=p1.size))
O4 = Oracle(synthetic.BALProblem(prob.cam_idx, prob.pt_idx, prob.obs, c1p.reshape(-1,9), p1p.reshape(-1,3), "y")); O4.set_robust("huber", 20.0, Pm)
_, sc4, _ = O4.linearize(); od4, _ = O4.solve(mu)
cams4 = c1p + od4[:9*prob.n_cams]*sc4[:9*prob.n_cams]; pts4 = p1p + od4[9*prob.n_cams:]*sc4[9*prob.n_cams:]
O5 = Oracle(synthetic.BALProblem(prob.cam_idx, prob.pt_idx, prob.obs, cams4.reshape(-1,9), pts4.reshape(-1,3), "z")); O5.set_robust("huber", 20.0, Pm)
_, c5 = O5.residuals()
print("oracle with parameters perturbed by 1e-13: cost %.10f rel to unperturbed %.2e" % (c5, abs(c5-c3)/c3))

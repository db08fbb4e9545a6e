"""Developer smoke script (GPU box): product vs oracle on small problems, stage by stage."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graphite_b200 import synthetic, binding
from oracle.binding import Oracle, default_options


def rel(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300))


def run(prob, precision="f64-f64", lm_iters=50):
    ctx = binding.Context(0)
    P = binding.problem_from_bal(ctx, prob, precision)
    print(prob.name, prob.shape(), precision, P.info())
    O = Oracle(prob, "f64" if precision.startswith("f64") else "f32")
    chi2 = P.linearize()
    ochi2, osc, ob = O.linearize()
    print(" chi2 %.17g oracle %.17g rel %.2e" % (chi2, ochi2, abs(chi2 - ochi2) / ochi2))
    print(" scales rel %.2e  b rel %.2e" % (rel(P.scales(), osc), rel(P.gradient(), ob)))
    r, _ = O.residuals()
    print(" residual rel %.2e" % rel(P.residuals(), r))
    O2 = Oracle(prob, "f64" if precision.startswith("f64") else "f32")
    ojc, ojp = O2.jacobians()
    jc, jp = P.jacobians()
    print(" Jc rel %.2e Jp rel %.2e" % (rel(jc, ojc), rel(jp, ojp)))
    cp, ri, off = P.hessian_structure()
    ocp, ori, ooff = O.hessian_structure()
    print(" structure equal", np.array_equal(cp, ocp), np.array_equal(ri, ori), np.array_equal(off, ooff))
    print(" H values rel %.2e" % rel(P.hessian_values(), O.hessian_values()))
    mu = 1e-4
    P.set_damping(mu)
    S, obS = O.schur(mu)
    print(" bS rel %.2e" % rel(P.schur_rhs(), obS))
    Sd = P.schur_diagonal()
    nc = prob.n_cams
    oSd = np.stack([S[9 * c:9 * c + 9, 9 * c:9 * c + 9] for c in range(nc)])
    print(" S diag blocks rel %.2e" % rel(Sd, oSd))
    Sfull = np.triu(S) + np.triu(S, 1).T
    rng = np.random.default_rng(0)
    xv = rng.normal(size=9 * nc)
    print(" S*x rel %.2e" % rel(P.schur_multiply(xv), Sfull @ xv))
    d, info = P.solve()
    od, ok = O.solve(mu)
    print(" solve: pcg iters", info, "oracle", ok, " delta rel %.2e (cams %.2e)" % (rel(d, od), rel(d[:9 * nc], od[:9 * nc])))
    new_chi2, rho_den = P.try_step()
    print(" try_step chi2 %.12g rho_den %.6g" % (new_chi2, rho_den))
    P.revert_step()
    print(" cost after revert %.17g" % P.compute_cost())
    # full LM
    P.set_vertices(prob.cams, prob.pts)
    t0 = time.time()
    traj, res = P.lm(iterations=lm_iters)
    dt = time.time() - t0
    O3 = Oracle(prob, "f64" if precision.startswith("f64") else "f32")
    otraj = O3.lm(default_options(iterations=lm_iters))
    n = min(len(traj), len(otraj))
    relc = np.abs(traj[:n, 1] - otraj[:n, 1]) / otraj[:n, 1]
    print(" LM: %d iters wall %.3fs device %.4fs; per-iter rel diff max %.2e final %.12g vs %.12g" % (
        len(traj), dt, res["seconds_total"], relc.max(), traj[-1, 1], otraj[-1, 1]))
    print("   pcg iters product", traj[:, 3].astype(int).tolist())
    print("   pcg iters oracle ", otraj[:, 3].astype(int).tolist())
    print("   stage seconds", {k: round(v, 5) for k, v in res.items() if k.startswith("seconds")})
    P.close(); ctx.close()


if __name__ == "__main__":
    which = sys.argv[1:] or ["fixture", "ladybug-49"]
    for w in which:
        name, _, prec = w.partition(":")
        prob = synthetic.schur_fixture() if name == "fixture" else synthetic.make_named(name)
        run(prob, prec or "f64-f64", 10 if name == "fixture" else 50)

"""Generate golden fixtures by running the UNMODIFIED reference on a B200.

TEST INFRASTRUCTURE.  Run on the GPU box (``gpurun -- python oracle/make_golden.py``):
it executes ``oracle/_ref/ref_bal`` (the reference's own GPU LM path compiled from
/root/reference by oracle/Makefile) on seeded synthetic problems and writes
``gpurun_out/golden/*.json|*.npz``; those files are then copied into
``tests/golden/`` and committed.  Nothing here is imported by the product.

What is recorded per case
  * trajectory: the (iteration, initial chi2, current chi2, lambda) table the
    reference prints (optimizer/levenberg_marquardt.hpp:216-221, 12 significant digits),
    final chi2 with 17 digits, wall seconds;
  * first linearisation (FP64 only, small cases): Hessian block-CSC structure, b,
    Jacobi scales, H values, Schur b_S and the Schur scalar upper CSC at lambda.
"""
from __future__ import annotations

import json
import os
import subprocess
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graphite_b200 import synthetic  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "ref_bal")
OUT = os.path.join(ROOT, "gpurun_out", "golden")


def parse_table(text: str):
    rows = []
    for line in text.splitlines():
        tok = line.split()
        if len(tok) == 6:
            try:
                rows.append([int(tok[0])] + [float(t) for t in tok[1:]])
            except ValueError:
                pass
    return rows


def run_case(name, prob, solver, precision, iterations, dump=False, lam=1e-4, timeout=3000, huber=0.0, weights=False, fix_cameras=0, fix_points=0):
    os.makedirs(OUT, exist_ok=True)
    gbal = os.path.join(OUT, f"{name}.gbal")
    if not os.path.exists(gbal):
        synthetic.write_gbal(prob, gbal)
    tag = f"{name}__{solver}__{precision}"
    cmd = [REF, gbal, "--solver", solver, "--precision", precision, "--iterations", str(iterations), "--lambda", repr(lam)]
    if huber > 0:  # HuberLoss(delta) on every factor (loss.hpp:27-51)
        tag += f"__huber{huber:g}"
        cmd += ["--huber", repr(huber)]
    if weights:    # per-factor precision matrices, synthetic.precision_matrices
        tag += "__weights"
        cmd += ["--weights"]
    if fix_cameras or fix_points:  # VertexDescriptor::set_fixed on the first cameras / every K-th point
        tag += f"__fixed{fix_cameras}c{fix_points}p"
        cmd += ["--fix_cameras", str(fix_cameras), "--fix_points", str(fix_points)]
    prefix = os.path.join(OUT, tag)
    if dump:
        cmd += ["--dump", prefix]
    t0 = time.time()
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
    wall = time.time() - t0
    rows = parse_table(res.stdout)
    final = [l for l in res.stdout.splitlines() if l.startswith("FINAL_CHI2")]
    total = [l for l in res.stdout.splitlines() if l.startswith("TOTAL_SECONDS")]
    dchi = [l for l in res.stdout.splitlines() if l.startswith("DUMP chi2")]
    rec = {
        "case": name, "shape": list(prob.shape()), "seed": 0, "solver": solver, "precision": precision,
        "lambda": lam, "huber": huber, "weights": weights, "fix_cameras": fix_cameras, "fix_points": fix_points, "pcg_iterations": 10, "pcg_tolerance": 1.0, "rejection_ratio": 5.0,
        "iterations": iterations, "returncode": res.returncode,
        "table_columns": ["iteration", "initial_chi2", "current_chi2", "lambda", "iter_seconds", "total_seconds"],
        "table": rows,
        "final_chi2": float(final[0].split()[1]) if final else None,
        "lm_seconds": float(total[0].split()[1]) if total else None,
        "initial_chi2_17g": float(dchi[0].split()[2]) if dchi else None,
        "wall_seconds": wall,
        "stderr_tail": res.stderr[-400:],
    }
    with open(prefix + ".json", "w") as fh:
        json.dump(rec, fh, indent=1)
    if dump and res.returncode == 0:
        T = np.float64 if precision.startswith("FP64") else np.float32
        arrs = {}
        for key, dt in [("H_colptr", np.uint64), ("H_rowidx", np.uint64), ("H_offsets", np.uint64), ("S_colptr", np.uint64),
                        ("S_rowidx", np.uint64), ("Scsc_ptr", np.int32), ("Scsc_idx", np.int32)]:
            ext = ".u64" if dt == np.uint64 else ".i32"
            a = np.fromfile(prefix + "." + key + ext, dtype=dt)
            arrs[key] = a.astype(np.int64 if dt == np.uint64 else np.int32)
            os.remove(prefix + "." + key + ext)
        for key in ["H_values", "b", "scales", "bS", "Scsc_val"]:
            arrs[key] = np.fromfile(prefix + "." + key + ".bin", dtype=T)
            os.remove(prefix + "." + key + ".bin")
        for key in ["final_cams", "final_pts"]:
            arrs[key] = np.fromfile(prefix + "." + key + ".f64", dtype=np.float64)
            os.remove(prefix + "." + key + ".f64")
        # keep fixtures small: H values are reduced to per-kind checksums plus the camera blocks
        nc = prob.n_cams
        hv = arrs.pop("H_values")
        arrs["H_cam_blocks"] = hv[: 81 * nc].copy()
        arrs["H_values_sum"] = np.array([hv.sum(dtype=np.float64), np.abs(hv).sum(dtype=np.float64), float(hv.size)])
        arrs["H_values_head"] = hv[81 * nc: 81 * nc + 4096].copy()
        np.savez_compressed(prefix + ".npz", **arrs)
    print(tag, "rc", res.returncode, "rows", len(rows), "final", rec["final_chi2"], f"lm {rec['lm_seconds']} wall {wall:.1f}s", flush=True)
    return rec


def reduce_schur(path, n_cams):
    """Keep a fixture small when S is nearly dense (long tracks connect almost every camera pair: 7 M scalars = 55 MB):
    replace the scalar upper CSC of S by what the tests compare - y = S x for the seeded vector the tests use, the diagonal
    blocks, and checksums (sum, sum of magnitudes, count) of the upper triangle.  `python oracle/make_golden.py --reduce F.npz N`."""
    z = dict(np.load(path))
    ptr, idx, val = z.pop("Scsc_ptr"), z.pop("Scsc_idx"), z.pop("Scsc_val")
    n = 9 * n_cams
    Sd = np.zeros((n, n))
    for c in range(n):
        Sd[idx[ptr[c]:ptr[c + 1]], c] = val[ptr[c]:ptr[c + 1]]
    Sfull = Sd + np.triu(Sd, 1).T
    x = np.random.default_rng(2).normal(size=n)
    z["S_times_x"] = Sfull @ x
    z["S_diag_blocks"] = np.stack([Sfull[9 * c:9 * c + 9, 9 * c:9 * c + 9] for c in range(n_cams)])
    z["S_sum"] = np.array([val.sum(dtype=np.float64), np.abs(val).sum(dtype=np.float64), float(val.size)])
    z["Scsc_ptr"] = ptr
    np.savez_compressed(path, **z)


def second_run(name, prob):
    """An extra FP64 run kept as *.run2.json: the reference's own run-to-run spread (float atomics)."""
    run_case(name, prob, "pcg-schur", "FP64-FP64", 50, timeout=6000)
    base = os.path.join(OUT, f"{name}__pcg-schur__FP64-FP64")
    os.replace(base + ".json", base + ".run2.json")


def main(which):
    if "fixture" in which:
        p = synthetic.schur_fixture()
        run_case("schur-fixture", p, "pcg-schur", "FP64-FP64", 10, dump=True)
    if "ladybug" in which:
        p = synthetic.make_named("ladybug-49")
        run_case("ladybug-49", p, "pcg-schur", "FP64-FP64", 50, dump=True)
        run_case("ladybug-49", p, "pcg-schur", "FP64-FP64", 50, dump=False)  # second run: run-to-run spread of the reference
        os.replace(os.path.join(OUT, "ladybug-49__pcg-schur__FP64-FP64.json"), os.path.join(OUT, "ladybug-49__pcg-schur__FP64-FP64.run2.json"))
        run_case("ladybug-49", p, "pcg-schur", "FP64-FP64", 50, dump=True)
        run_case("ladybug-49", p, "pcg-schur", "FP32-FP32", 50)
        run_case("ladybug-49", p, "pcg", "FP64-FP32", 50)
        run_case("ladybug-49", p, "pcg", "FP64-FP64", 50)
    if "trafalgar" in which:
        p = synthetic.make_named("trafalgar-257")
        second_run("trafalgar-257", p)
        run_case("trafalgar-257", p, "pcg-schur", "FP64-FP64", 50)
        run_case("trafalgar-257", p, "pcg-schur", "FP32-FP32", 50)
        run_case("trafalgar-257", p, "pcg", "FP64-FP32", 50)
    if "robust" in which:
        p = synthetic.make_named("ladybug-49")
        run_case("ladybug-49", p, "pcg-schur", "FP64-FP64", 50, dump=True, huber=20.0, weights=True)
        run_case("ladybug-49", p, "pcg-schur", "FP64-FP64", 50, huber=20.0)
        run_case("ladybug-49", p, "pcg-schur", "FP64-FP64", 50, weights=True)
        run_case("ladybug-49", p, "pcg", "FP64-FP64", 50, huber=20.0, weights=True)
        p = synthetic.make_named("trafalgar-257")
        run_case("trafalgar-257", p, "pcg-schur", "FP64-FP64", 50, huber=20.0, weights=True)
    if "robust2" in which:  # the reference's own run-to-run spread on the robust cases (float atomics)
        for name, kw in [("ladybug-49", dict(huber=20.0, weights=True)), ("trafalgar-257", dict(huber=20.0, weights=True))]:
            p = synthetic.make_named(name)
            rec = run_case(name, p, "pcg-schur", "FP64-FP64", 50, **kw)
            base = os.path.join(OUT, f"{name}__pcg-schur__FP64-FP64__huber20__weights")
            os.replace(base + ".json", base + ".run2.json")
    if "fixed" in which:  # fixed vertices (gauge cameras, some fixed points): pcg-schur and pcg, final vertices dumped
        p = synthetic.make_named("ladybug-49")
        run_case("ladybug-49", p, "pcg-schur", "FP64-FP64", 30, dump=True, fix_cameras=2, fix_points=50)
        run_case("ladybug-49", p, "pcg", "FP64-FP64", 30, fix_cameras=2, fix_points=50)
        run_case("ladybug-49", p, "pcg-schur", "FP32-FP32", 30, fix_cameras=2, fix_points=50)
        p = synthetic.make_named("trafalgar-257")
        run_case("trafalgar-257", p, "pcg-schur", "FP64-FP64", 30, fix_cameras=1, fix_points=0)
    if "longtracks" in which:  # tracks longer than one tile of the library holds (400 / 260 / 193 / 300 cameras)
        p = synthetic.make_named("long-tracks")
        run_case("long-tracks", p, "pcg-schur", "FP64-FP64", 30, dump=True)
        reduce_schur(os.path.join(OUT, "long-tracks__pcg-schur__FP64-FP64.npz"), p.n_cams)
        run_case("long-tracks", p, "pcg", "FP64-FP64", 30)
        run_case("long-tracks", p, "pcg-schur", "FP32-FP32", 30)
        run_case("long-tracks", p, "pcg", "FP64-FP32", 30)
        run_case("long-tracks", p, "pcg", "FP64-BF16", 30)
    if "dubrovnik" in which:
        p = synthetic.make_named("dubrovnik-356")
        second_run("dubrovnik-356", p)
        run_case("dubrovnik-356", p, "pcg-schur", "FP64-FP64", 50)
    if "venice" in which:
        p = synthetic.make_named("venice-1778")
        second_run("venice-1778", p)
        run_case("venice-1778", p, "pcg-schur", "FP64-FP64", 50, timeout=6000)
    if "mixed" in which:  # round 2: the precision matrix of examples/bal.cu:159-236 at the sizes round 1 left out
        p = synthetic.make_named("venice-1778")
        run_case("venice-1778", p, "pcg-schur", "FP32-FP32", 50, timeout=6000)
        run_case("venice-1778", p, "pcg", "FP64-FP32", 50, timeout=6000)
        run_case("venice-1778", p, "pcg", "FP64-BF16", 50, timeout=6000)
        for name in ("ladybug-49", "trafalgar-257"):
            p = synthetic.make_named(name)
            run_case(name, p, "pcg", "FP64-BF16", 50)
        p = synthetic.make_named("dubrovnik-356")
        run_case("dubrovnik-356", p, "pcg-schur", "FP32-FP32", 50)
    if "final" in which:  # BASELINE configs[4]: Final-13682 shape, FP32 / mixed
        p = synthetic.make_named("final-13682")
        run_case("final-13682", p, "pcg-schur", "FP32-FP32", 50, timeout=6000)
        run_case("final-13682", p, "pcg", "FP64-FP32", 30, timeout=6000)
    # .gbal inputs are regenerated from the seed; do not ship them back
    for f in os.listdir(OUT):
        if f.endswith(".gbal"):
            os.remove(os.path.join(OUT, f))


if __name__ == "__main__":
    if len(sys.argv) == 4 and sys.argv[1] == "--reduce":
        reduce_schur(sys.argv[2], int(sys.argv[3]))
    else:
        main(sys.argv[1:] or ["fixture", "ladybug", "trafalgar"])

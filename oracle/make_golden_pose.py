"""Generate the pose-graph golden fixtures by running the UNMODIFIED reference on a B200.

TEST INFRASTRUCTURE.  Run on the GPU box (``gpurun -- python oracle/make_golden_pose.py``): executes
``oracle/_ref/ref_pose`` (oracle/ref_pose_driver.cu: the reference's generic factor machinery, PCGSolver +
BlockJacobiPreconditioner, levenberg_marquardt) on ``synthetic.pose_graph()`` and writes
``gpurun_out/golden/pose-graph*.json|npz``, which are then copied into ``tests/golden/`` and committed.
"""
from __future__ import annotations

import json
import os
import subprocess
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graphite_b200 import synthetic  # noqa: E402
from oracle.make_golden import parse_table  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "ref_pose")
OUT = os.path.join(ROOT, "gpurun_out", "golden")

PROTOCOL = dict(lam=1e-4, iterations=12, pcg_iterations=30, pcg_tolerance=1e-10, rejection_ratio=5.0)


def run_case(tag, pg, precision, level=0, dump=True, **kw):
    o = dict(PROTOCOL, **kw)
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, "pose-graph.gpg")
    synthetic.write_pose_graph(pg, path)
    prefix = os.path.join(OUT, f"{tag}__pcg__{precision}__level{level}")
    cmd = [REF, path, "--precision", precision, "--level", str(level), "--lambda", repr(o["lam"]), "--iterations", str(o["iterations"]),
           "--pcg_iterations", str(o["pcg_iterations"]), "--pcg_tolerance", repr(o["pcg_tolerance"]), "--rejection_ratio", repr(o["rejection_ratio"])]
    if dump:
        cmd += ["--dump", prefix]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=1200)
    lines = res.stdout.splitlines()
    final = [l for l in lines if l.startswith("FINAL_CHI2")]
    dchi = [l for l in lines if l.startswith("DUMP chi2")]
    ddim = [l for l in lines if l.startswith("DUMP dimH")]
    rec = dict(case=tag, precision=precision, level=level, protocol=o, returncode=res.returncode, table=parse_table(res.stdout),
               table_columns=["iteration", "initial_chi2", "current_chi2", "lambda", "iter_seconds", "total_seconds"],
               final_chi2=float(final[0].split()[1]) if final else None, initial_chi2_17g=float(dchi[0].split()[2]) if dchi else None,
               hessian_dim=int(ddim[0].split()[2]) if ddim else None, rejected_pcg_updates=res.stdout.count("rejected pcg update"),
               stderr_tail=res.stderr[-400:])
    with open(prefix + ".json", "w") as fh:
        json.dump(rec, fh, indent=1)
    if dump and res.returncode == 0:
        T = np.float64 if precision.startswith("FP64") else np.float32
        S = np.float64 if precision.endswith("FP64") else np.float32
        arrs = {}
        for key in ("H_colptr", "H_rowidx", "H_offsets"):
            arrs[key] = np.fromfile(f"{prefix}.{key}.u64", dtype=np.uint64).astype(np.int64)
            os.remove(f"{prefix}.{key}.u64")
        arrs["H_values"] = np.fromfile(prefix + ".H_values.bin", dtype=S)
        arrs["b"] = np.fromfile(prefix + ".b.bin", dtype=T)
        arrs["scales"] = np.fromfile(prefix + ".scales.bin", dtype=T)
        arrs["columns"] = np.fromfile(prefix + ".columns.i64", dtype=np.int64)
        arrs["final_poses"] = np.fromfile(prefix + ".final_poses.f64", dtype=np.float64)
        for ext in (".H_values.bin", ".b.bin", ".scales.bin", ".columns.i64", ".final_poses.f64"):
            os.remove(prefix + ext)
        np.savez_compressed(prefix + ".npz", **arrs)
    print(os.path.basename(prefix), "rc", res.returncode, "rows", len(rec["table"]), "final", rec["final_chi2"], flush=True)
    if res.returncode != 0:
        print(res.stdout[-2000:], res.stderr[-2000:])
    os.remove(path)
    return rec


def main():
    pg = synthetic.pose_graph()
    run_case("pose-graph", pg, "FP64-FP64", level=0)
    run_case("pose-graph", pg, "FP64-FP64", level=1)
    run_case("pose-graph", pg, "FP32-FP32", level=0)
    run_case("pose-graph", pg, "FP64-FP32", level=0)
    # a run with rejected steps: large initial perturbation (tests/test_gpu_graph.py: PROTO_HARD)
    run_case("pose-graph-hard", synthetic.pose_graph_hard(), "FP64-FP64", level=0, lam=1e-4, iterations=14, pcg_iterations=100)


if __name__ == "__main__":
    main()

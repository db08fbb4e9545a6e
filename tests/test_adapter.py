"""GPU (-m gpu): the compiled reference-side adapter (include/graphite_b200_adapter.hpp).

oracle/_ref/adapter_test is built HERE from the unmodified reference headers under /root/reference/include plus the adapter
header (oracle/Makefile; the binary travels to the GPU box like the library).  It runs the reference's own
optimizer::levenberg_marquardt three ways on one problem - with its PCGSchurSolver, with graphite::B200SchurSolver plugged
in as the Solver<T,S>, and the library's whole loop over the real descriptors - and one solve of the first linearisation
with both solvers (the cross-solver check of tests/schur.cu:340-389, bound 5e-4 there)."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from graphite_b200 import synthetic

pytestmark = pytest.mark.gpu
EXE = os.path.join(ROOT, "oracle", "_ref", "adapter_test")


def run_adapter(tmp_path, case, *extra):
    if not os.path.exists(EXE):
        pytest.fail("oracle/_ref/adapter_test is missing: run `python -c 'import __graft_entry__ as g; g.build()'` where "
                    "/root/reference exists")
    prob = synthetic.schur_fixture() if case == "schur-fixture" else synthetic.make_named(case)
    path = str(tmp_path / f"{case}.gbal")
    synthetic.write_gbal(prob, path)
    res = subprocess.run([EXE, path, *extra], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    sections, cur, info = {}, None, {}
    for line in res.stdout.splitlines():
        tok = line.split()
        if line.startswith("== "):
            cur = tok[1]
            sections[cur] = {"rows": []}
        elif line.startswith("DELTA_REL"):
            info["delta_rel"] = float(tok[1])
            info["pcg_iterations"] = int(tok[5])
        elif cur and tok and tok[0] in ("FINAL_CHI2", "LIB_FINAL_CHI2", "SECONDS"):
            sections[cur][tok[0]] = float(tok[1])
        elif cur and tok and tok[0] == "OK":
            sections[cur]["OK"] = int(tok[1])
        elif cur and len(tok) == 6:
            try:
                sections[cur]["rows"].append([float(tok[1]), float(tok[2]), float(tok[3])])  # initial, current, lambda
            except ValueError:
                pass
    for k in sections:
        sections[k]["rows"] = np.array(sections[k]["rows"])
    return info, sections, res.stdout


@pytest.mark.parametrize("case,iters", [("schur-fixture", 10), ("ladybug-49", 30), ("trafalgar-257", 25)])
def test_b200_solver_plugs_into_the_reference_lm(built, tmp_path, case, iters):
    info, sec, out = run_adapter(tmp_path, case, "--iterations", str(iters))
    # one solve of the same linearisation: same PCG (10 iterations), so far tighter than the reference's 5e-4
    assert info["delta_rel"] <= 1e-7, out[-1500:]
    ref, mine, loop = sec["REFERENCE"], sec["B200SOLVER"], sec["B200LOOP"]
    assert len(ref["rows"]) == len(mine["rows"]) == iters
    r = np.abs(mine["rows"][:, 1] - ref["rows"][:, 1]) / ref["rows"][:, 1]
    assert r.max() <= 5e-9, r  # the table prints 12 digits; the reference's own run-to-run spread is ~1e-10
    assert np.array_equal(mine["rows"][:, 0] == mine["rows"][:, 1], ref["rows"][:, 0] == ref["rows"][:, 1]), "decisions differ"
    np.testing.assert_allclose(mine["rows"][:, 2], ref["rows"][:, 2], rtol=1e-6)
    assert abs(mine["FINAL_CHI2"] - ref["FINAL_CHI2"]) <= 1e-6 * ref["FINAL_CHI2"]
    # the whole loop: same trajectory, and the vertices the USER holds (updated in place through VertexTraits::update)
    # have the cost the library reports, evaluated by the reference's own kernels
    assert loop["OK"] == 1
    assert len(loop["rows"]) == iters
    r2 = np.abs(loop["rows"][:, 1] - ref["rows"][:, 1]) / ref["rows"][:, 1]
    assert r2.max() <= 5e-9, r2
    assert abs(loop["FINAL_CHI2"] - loop["LIB_FINAL_CHI2"]) <= 1e-9 * loop["LIB_FINAL_CHI2"]
    assert abs(loop["FINAL_CHI2"] - ref["FINAL_CHI2"]) <= 1e-6 * ref["FINAL_CHI2"]


def test_b200_solver_with_huber_loss_and_precision_matrices(built, tmp_path):
    """The descriptor's HuberLoss and per-factor precision matrices reach the library through the adapter."""
    info, sec, out = run_adapter(tmp_path, "ladybug-49", "--iterations", "12", "--huber", "20", "--weights")
    assert info["delta_rel"] <= 1e-7, out[-1500:]
    ref, mine = sec["REFERENCE"], sec["B200SOLVER"]
    r = np.abs(mine["rows"][:, 1] - ref["rows"][:, 1]) / ref["rows"][:, 1]
    assert r.max() <= 5e-9, r
    assert np.array_equal(mine["rows"][:, 0] == mine["rows"][:, 1], ref["rows"][:, 0] == ref["rows"][:, 1])


def test_b200_solver_fp32(built, tmp_path):
    info, sec, out = run_adapter(tmp_path, "ladybug-49", "--iterations", "15", "--precision", "FP32-FP32")
    assert info["delta_rel"] <= 5e-4, out[-1500:]  # tests/schur.cu:386
    ref, mine = sec["REFERENCE"], sec["B200SOLVER"]
    r = np.abs(mine["rows"][:, 1] - ref["rows"][:, 1]) / ref["rows"][:, 1]
    assert r.max() <= 1e-4, r


# ------------------------------------------------------------------------------------------------------
# generic graphs: graphite::B200GraphSolver (include/graphite_b200_graph_adapter.hpp) on the pose-graph fixture
# ------------------------------------------------------------------------------------------------------
GEXE = os.path.join(ROOT, "oracle", "_ref", "adapter_graph_test")


def run_graph_adapter(tmp_path, pg, *extra):
    if not os.path.exists(GEXE):
        pytest.fail("oracle/_ref/adapter_graph_test is missing: run `python -c 'import __graft_entry__ as g; g.build()'` where "
                    "/root/reference exists")
    path = str(tmp_path / "pose.gpg")
    synthetic.write_pose_graph(pg, path)
    res = subprocess.run([GEXE, path, *extra], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    sections, cur, info = {}, None, {}
    for line in res.stdout.splitlines():
        tok = line.split()
        if line.startswith("== "):
            cur = tok[1]
            sections[cur] = {"rows": []}
        elif line.startswith("DELTA_REL"):
            info["delta_rel"] = float(tok[1])
            info["pcg_iterations"] = int(tok[5])
        elif cur and tok and tok[0] == "FINAL_CHI2":
            sections[cur]["FINAL_CHI2"] = float(tok[1])
        elif cur and len(tok) == 6:
            try:
                sections[cur]["rows"].append([float(tok[1]), float(tok[2]), float(tok[3])])
            except ValueError:
                pass
    for k in sections:
        sections[k]["rows"] = np.array(sections[k]["rows"])
    return info, sections, res.stdout


@pytest.mark.parametrize("level", [0, 1])
def test_b200_graph_solver_plugs_into_the_reference_lm(built, tmp_path, level):
    """ANY Graphite graph: the reference linearises the user's pose-graph factors (autodiff edges with Huber loss, precision
    matrices and activity levels; unary priors; fixed vertices) and B200GraphSolver solves on its buffers in place."""
    info, sec, out = run_graph_adapter(tmp_path, synthetic.pose_graph(), "--level", str(level))
    assert info["delta_rel"] <= 1e-7, out[-1500:]
    ref, mine = sec["REFERENCE"], sec["B200SOLVER"]
    assert len(ref["rows"]) == len(mine["rows"]) == 12
    r = np.abs(mine["rows"][:, 1] - ref["rows"][:, 1]) / ref["rows"][0, 0]
    assert r.max() <= 5e-9, r
    assert np.array_equal(mine["rows"][:, 0] == mine["rows"][:, 1], ref["rows"][:, 0] == ref["rows"][:, 1]), "decisions differ"
    np.testing.assert_allclose(mine["rows"][:, 2], ref["rows"][:, 2], rtol=1e-6)
    assert abs(mine["FINAL_CHI2"] - ref["FINAL_CHI2"]) <= 1e-6 * ref["FINAL_CHI2"]


def test_b200_graph_solver_with_rejected_steps(built, tmp_path):
    info, sec, out = run_graph_adapter(tmp_path, synthetic.pose_graph_hard(), "--iterations", "14", "--pcg_iterations", "100")
    ref, mine = sec["REFERENCE"], sec["B200SOLVER"]
    assert (ref["rows"][:, 0] == ref["rows"][:, 1]).any(), "the fixture is meant to reject steps"
    assert np.array_equal(mine["rows"][:, 0] == mine["rows"][:, 1], ref["rows"][:, 0] == ref["rows"][:, 1]), "decisions differ"
    r = np.abs(mine["rows"][:, 1] - ref["rows"][:, 1]) / ref["rows"][0, 0]
    assert r.max() <= 1e-8, r  # far-from-optimum start: rounding is amplified ~1000x (tests/test_gpu_graph.py)


@pytest.mark.parametrize("precision,tol", [("FP32-FP32", 1e-4), ("FP64-FP32", 1e-4)])
def test_b200_graph_solver_low_precision(built, tmp_path, precision, tol):
    info, sec, out = run_graph_adapter(tmp_path, synthetic.pose_graph(), "--precision", precision)
    assert info["delta_rel"] <= 5e-4, out[-1500:]
    ref, mine = sec["REFERENCE"], sec["B200SOLVER"]
    r = np.abs(mine["rows"][:, 1] - ref["rows"][:, 1]) / ref["rows"][0, 0]
    assert r.max() <= tol, r

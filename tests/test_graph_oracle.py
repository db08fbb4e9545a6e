"""CPU: the numpy oracle of the generic factor-graph path (oracle/oracle_graph.py) against the integer known-answer tests
of the reference's tests/factor.cu and against runs of the unmodified reference on the pose-graph fixture
(oracle/ref_pose_driver.cu -> tests/golden/pose-graph*)."""
import numpy as np
import pytest

from conftest import golden_json, golden_npz
from graphite_b200 import synthetic
from oracle import oracle_graph as og

UNARY = [np.array([[1.0, 0.0]])]
COUPLED = [np.array([[2.0, 3.0]])]
BINARY = [np.array([[1.0, 2.0]]), np.array([[3.0, 4.0]])]


def toy(factors, vertices, fixed=None):
    O = og.GraphOracle()
    O.scale_on = False
    ids = [v[0] for v in vertices]
    vs = O.add_vertex_set(2, ids, [v[1] for v in vertices], fixed)
    for A, conn, obs, loss, delta, active in factors:
        idx = [[ids.index(g) for g in c] for c in conn]
        O.add_factor_set(A[0].shape[0], [vs] * len(A), idx, og.make_linear(A, np.asarray(obs, dtype=np.float64).reshape(len(conn), -1)),
                         active=active, loss=loss, delta=delta)
    return O


def test_kats_of_the_reference_factor_tests():
    # ComputeB (tests/factor.cu:425-466): two identical unary factors, r = 4.5 each -> b0 = -9 per call
    O = toy([(UNARY, [[10], [10]], [2.5, 2.5], 0, 0.0, None)], [(10, (7.0, 0.0))])
    O.initialize(0)
    assert O.linearize() == 40.5 and np.array_equal(O.b, [-9.0, 0.0])
    # ComputeBHuberLoss (:468-509): each factor contributes -1; Chi2HuberLoss (:758-784): 8 + 0.25
    O = toy([(UNARY, [[10], [10]], [2.5, 2.5], 1, 1.0, None)], [(10, (7.0, 0.0))])
    O.initialize(0); O.linearize()
    assert np.allclose(O.b, [-2.0, 0.0], rtol=1e-15)
    assert np.array_equal(O.jv(np.array([3.0, 5.0])), [3.0, 3.0])            # ComputeJvHuberLoss (:597-640)
    assert np.allclose(O.jtpv(np.array([9.0, 9.0])), [4.0, 0.0], rtol=1e-15)  # ComputeJtvHuberLoss (:677-720)
    O = toy([(UNARY, [[10], [10]], [2.5, 6.5], 1, 1.0, None)], [(10, (7.0, 0.0))])
    O.initialize(0)
    assert O.linearize() == 8.25 and np.array_equal(O.F[0]["chi2"], [8.0, 0.25])
    # ComputeHessianBlockDiagonal (:511-555) [8,12,12,18], ComputeHessianScalarDiagonal (:557-595) [8,18]
    O = toy([(COUPLED, [[10], [10]], [2.5, 2.5], 0, 0.0, None)], [(10, (7.0, 0.0))])
    O.initialize(0); O.linearize()
    assert np.array_equal(O.H.T.reshape(-1), [8, 12, 12, 18]) and np.array_equal(np.diag(O.H), [8, 18])
    # ComputeHessian (:854-967)
    O = toy([(COUPLED, [[10], [20]], [2.5, 3.5], 0, 0.0, None), (BINARY, [[10, 20]], [4.5], 0, 0.0, None)], [(10, (7.0, 5.0)), (20, (11.0, 13.0))])
    O.initialize(0); O.linearize()
    assert O.colptr.tolist() == [0, 1, 3] and O.rowidx.tolist() == [0, 0, 1] and O.offsets.tolist() == [0, 4, 8]
    assert np.array_equal(O.hessian_values(), [5, 8, 8, 13, 3, 6, 4, 8, 13, 18, 18, 25])


def test_fixed_and_level_semantics():
    verts = [(10, (7.0, 5.0)), (20, (11.0, 13.0)), (30, (1.0, 2.0))]
    facs = [(COUPLED, [[10], [20], [30]], [2.5, 3.5, 0.5], 0, 0.0, [0, 0, 1]), (BINARY, [[10, 20]], [4.5], 0, 0.0, None)]
    O = toy(facs, verts, fixed=[0, 1, 0])
    assert O.initialize(0) == 2 and O.V[0]["hoff"].tolist() == [0, -1, -1]
    assert O.initialize(1) == 4 and O.V[0]["hoff"].tolist() == [0, -1, 2]
    O = toy([(UNARY, [[10], [20]], [2.5, 3.5], 0, 0.0, [0, 0x80])], verts[:2])  # disabled by the top bit at every level
    assert O.initialize(127) == 2


def pose_oracle(pg, level=0):
    O = og.GraphOracle()
    v = O.add_vertex_set(6, pg.ids, pg.poses, pg.fixed)
    O.add_factor_set(6, [v, v], pg.bt_idx, og.make_between6(pg.bt_meas), active=pg.bt_active, P=pg.bt_P, loss=1, delta=pg.huber)
    O.add_factor_set(6, [v], pg.pr_idx, og.make_prior6(pg.pr_meas))
    O.initialize(level)
    return O


def test_dual_number_jacobians_match_finite_differences():
    pg = synthetic.pose_graph()
    ev = og.make_between6(pg.bt_meas)
    xi, xj = pg.poses[2].copy(), pg.poses[7].copy()
    r, (Ji, Jj) = ev([xi, xj], 3)
    h = 1e-6
    for k in range(6):
        d = np.zeros(6); d[k] = h
        fi = (np.array(og.between6_residual(list(xi + d), list(xj), pg.bt_meas[3])) - np.array(og.between6_residual(list(xi - d), list(xj), pg.bt_meas[3]))) / (2 * h)
        fj = (np.array(og.between6_residual(list(xi), list(xj + d), pg.bt_meas[3])) - np.array(og.between6_residual(list(xi), list(xj - d), pg.bt_meas[3]))) / (2 * h)
        assert np.abs(fi - Ji[:, k]).max() < 1e-7 and np.abs(fj - Jj[:, k]).max() < 1e-7


@pytest.mark.parametrize("level", [0, 1])
def test_pose_graph_oracle_matches_reference(level):
    """Pins the oracle: structure bit-exact, first linearisation 1e-12, trajectory 1e-9 against the unmodified reference."""
    name = f"pose-graph__pcg__FP64-FP64__level{level}"
    try:
        js, z = golden_json(name + ".json"), golden_npz(name + ".npz")
    except FileNotFoundError:
        pytest.skip("golden not generated yet")
    O = pose_oracle(synthetic.pose_graph(), level)
    assert O.dimH == js["hessian_dim"]
    assert np.array_equal(O.colptr, z["H_colptr"]) and np.array_equal(O.rowidx, z["H_rowidx"]) and np.array_equal(O.offsets, z["H_offsets"])
    assert np.array_equal(O.V[0]["hoff"], z["columns"])
    chi2 = O.linearize()
    assert abs(chi2 - js["initial_chi2_17g"]) / chi2 < 1e-12
    r = lambda a, b: np.abs(a - b).max() / np.abs(b).max()
    assert r(O.scales, z["scales"]) < 1e-12 and r(O.b, z["b"]) < 1e-11 and r(O.hessian_values(), z["H_values"]) < 1e-11
    p = js["protocol"]
    tab = O.lm(iterations=p["iterations"], initial_damping=p["lam"], pcg_iterations=p["pcg_iterations"], pcg_tolerance=p["pcg_tolerance"],
               rejection_ratio=p["rejection_ratio"])
    ref = np.array(js["table"])
    assert np.array_equal(tab[:, 0] == tab[:, 1], ref[:, 1] == ref[:, 2])
    assert np.abs(tab[:, 1] - ref[:, 2]).max() / ref[0, 1] < 1e-9
    assert abs(tab[-1, 1] - js["final_chi2"]) / js["final_chi2"] < 1e-6

"""Deterministic synthetic BAL-shaped problems.

The reference ships no datasets and no generator; ``examples/bal.cu:63-147`` only
parses the BAL text format (``<n_cams> <n_pts> <n_obs>``, then one
``cam pt x y`` line per observation, then 9 values per camera ``[w(3) t(3) f k1 k2]``
and 3 per point).  This module emits problems of the named BAL shapes
(BASELINE.json ``configs``; counts from SURVEY.md section 8) in exactly that
convention: camera vertex id = camera index, point vertex id = ``n_cams + index``
(``examples/bal.cu:108,124,141``), observations sorted by (point, camera).

Scene: a "street" -- cameras along the x axis looking down -z (BAL cameras look
down the negative z axis, ``p = -P.xy / P.z``, ``examples/reprojection_error.cuh:82``),
points in a slab in front of them.  Each point is seen by >= 2 distinct cameras
drawn from the cameras that have it inside a +-27 degree field of view, which gives
the banded camera-camera coupling real BAL sequences show.  Observations carry 0.5 px
noise plus 5 % gross outliers; the initial state is the ground truth perturbed by 1e-2
(camera pose) and 3 % of depth (points).  numpy's PCG64 stream is
bit-reproducible across machines, so seed 0 here and on the GPU box agree.
"""
from __future__ import annotations

import dataclasses
import struct

import numpy as np

SHAPES = {
    # name: (n_cams, n_pts, n_obs)  -- SURVEY.md section 8 table
    "ladybug-49": (49, 7_776, 31_843),
    "trafalgar-257": (257, 65_132, 225_911),
    "dubrovnik-356": (356, 226_730, 1_255_268),
    "venice-1778": (1_778, 993_923, 5_001_946),
    "final-13682": (13_682, 4_456_117, 28_987_644),
}


@dataclasses.dataclass
class BALProblem:
    cam_idx: np.ndarray  # int32 [M]
    pt_idx: np.ndarray  # int32 [M], non-decreasing
    obs: np.ndarray  # float64 [M, 2]
    cams: np.ndarray  # float64 [Nc, 9] initial estimate
    pts: np.ndarray  # float64 [Np, 3] initial estimate
    name: str = "custom"

    @property
    def n_cams(self) -> int:
        return int(self.cams.shape[0])

    @property
    def n_pts(self) -> int:
        return int(self.pts.shape[0])

    @property
    def n_obs(self) -> int:
        return int(self.cam_idx.shape[0])

    def shape(self):
        return (self.n_cams, self.n_pts, self.n_obs)


def _rodrigues(w: np.ndarray) -> np.ndarray:
    """Rotation matrices [N,3,3] from angle-axis vectors [N,3]."""
    theta = np.linalg.norm(w, axis=1)
    a = w / theta[:, None]
    c, s = np.cos(theta), np.sin(theta)
    K = np.zeros((w.shape[0], 3, 3))
    K[:, 0, 1], K[:, 0, 2] = -a[:, 2], a[:, 1]
    K[:, 1, 0], K[:, 1, 2] = a[:, 2], -a[:, 0]
    K[:, 2, 0], K[:, 2, 1] = -a[:, 1], a[:, 0]
    eye = np.eye(3)[None]
    return c[:, None, None] * eye + s[:, None, None] * K + (1 - c)[:, None, None] * (a[:, :, None] * a[:, None, :])


def project(cams: np.ndarray, pts: np.ndarray, cam_idx: np.ndarray, pt_idx: np.ndarray) -> np.ndarray:
    """BAL projection f*(1+k1 r^2+k2 r^4)*p of each (camera, point) pair; float64 [M,2]."""
    R = _rodrigues(cams[:, :3])
    P = np.einsum("mij,mj->mi", R[cam_idx], pts[pt_idx]) + cams[cam_idx, 3:6]
    p = -P[:, :2] / P[:, 2:3]
    r2 = (p * p).sum(1)
    c = cams[cam_idx]
    return (c[:, 6] * (1.0 + c[:, 7] * r2 + c[:, 8] * r2 * r2))[:, None] * p


def make_bal(n_cams: int, n_pts: int, n_obs: int, seed: int = 0, name: str = "custom",
             noise_px: float = 0.5, cam_sigma: float = 1e-2, pt_sigma: float = 3e-2,
             outlier_frac: float = 0.05, outlier_px: float = 50.0) -> BALProblem:
    if n_obs < 2 * n_pts:
        raise ValueError("need at least two observations per point")
    rng = np.random.Generator(np.random.PCG64(seed))
    spacing = 0.25
    length = spacing * n_cams
    # --- ground-truth cameras ---------------------------------------------------------
    x0 = -0.5 * length  # scene centred on the origin
    centre = np.stack([x0 + spacing * (np.arange(n_cams) + 0.5), rng.normal(0, 0.3, n_cams), rng.normal(0, 0.3, n_cams)], 1)
    w = rng.normal(0, 0.08, (n_cams, 3))
    w[np.linalg.norm(w, axis=1) < 1e-3] = [0.01, -0.02, 0.015]  # keep theta > 0 (reprojection_error.cuh:72)
    R = _rodrigues(w)
    t = -np.einsum("nij,nj->ni", R, centre)
    f = rng.uniform(400.0, 1200.0, n_cams)
    k1 = rng.normal(0, 0.05, n_cams)
    k2 = rng.normal(0, 0.005, n_cams)
    cams_gt = np.concatenate([w, t, f[:, None], k1[:, None], k2[:, None]], 1)
    # --- ground-truth points (index grows with x: incremental-SfM-like locality) ---------
    x = np.sort(rng.uniform(x0, x0 + length, n_pts))
    depth = rng.uniform(5.0, 50.0, n_pts)
    y = rng.uniform(-0.25, 0.25, n_pts) * depth
    pts_gt = np.stack([x, y, -depth], 1)
    # --- visibility -------------------------------------------------------------------------
    half = 0.5 * depth
    lo = np.clip(np.ceil((x - x0 - half) / spacing - 0.5).astype(np.int64), 0, n_cams - 1)
    hi = np.clip(np.floor((x - x0 + half) / spacing - 0.5).astype(np.int64), 0, n_cams - 1)
    ncand = np.maximum(hi - lo + 1, 1)
    if (ncand < 2).any():
        raise ValueError("scene too small: a point has fewer than two candidate cameras")
    extra_mean = n_obs / n_pts - 2.0
    tl = 2 + rng.geometric(1.0 / (1.0 + extra_mean), n_pts) - 1
    tl = np.minimum(tl, ncand)
    # hit n_obs exactly, deterministically
    diff = int(n_obs - tl.sum())
    order = rng.permutation(n_pts)
    guard = 0
    while diff != 0:
        guard += 1
        if guard > 1000:
            raise ValueError("cannot reach the requested observation count")
        if diff > 0:
            ok = order[tl[order] < ncand[order]]
            take = ok[: min(diff, ok.size)]
            tl[take] += 1
            diff -= take.size
        else:
            ok = order[tl[order] > 2]
            take = ok[: min(-diff, ok.size)]
            tl[take] -= 1
            diff += take.size
        order = np.roll(order, 7919)
    pt_idx = np.repeat(np.arange(n_pts, dtype=np.int64), tl)
    start = np.cumsum(tl) - tl
    k = np.arange(n_obs, dtype=np.int64) - np.repeat(start, tl)  # rank within the track
    u = rng.random(n_obs)
    tl_o, nc_o, lo_o = tl[pt_idx], ncand[pt_idx], lo[pt_idx]
    # stratified draw: stratum k of width nc/tl >= 1 => distinct, ascending cameras per point
    cam_idx = lo_o + np.minimum(np.floor((k + u) * nc_o / tl_o).astype(np.int64), nc_o - 1)
    # strata of width exactly 1 can collide after the min(); repair by forcing strict ascent
    same = (np.diff(cam_idx, prepend=-1) <= 0) & (k > 0)
    while same.any():
        cam_idx[same] += 1
        same = (np.diff(cam_idx, prepend=-1) <= 0) & (k > 0)
    if cam_idx.max() >= n_cams:
        over = cam_idx >= n_cams
        raise ValueError(f"visibility repair overflowed for {int(over.sum())} observations")
    unseen = np.setdiff1d(np.arange(n_cams), cam_idx)
    if unseen.size:
        raise ValueError(f"{unseen.size} cameras observe no point; use more observations")
    cam_idx = cam_idx.astype(np.int32)
    pt_idx = pt_idx.astype(np.int32)
    # --- observations and perturbed initial state ----------------------------------------------
    obs = project(cams_gt, pts_gt, cam_idx, pt_idx) + rng.normal(0, noise_px, (n_obs, 2))
    # gross outliers (mismatched features), as real BAL data has: keeps the cost floor high, so LM
    # rejects steps and PCG runs several iterations instead of converging in two steps
    bad = rng.random(n_obs) < outlier_frac
    obs[bad] += rng.normal(0, outlier_px, (int(bad.sum()), 2))
    # perturb rotation and camera CENTRE (not t = -R c: with the origin hundreds of units away a 1e-2 rad
    # change of R at fixed t would move the camera by several units)
    w0 = w + rng.normal(0, cam_sigma, (n_cams, 3))
    c0 = centre + rng.normal(0, cam_sigma, (n_cams, 3))
    t0 = -np.einsum("nij,nj->ni", _rodrigues(w0), c0)
    cams = np.concatenate([w0, t0, (f * (1.0 + rng.normal(0, 1e-3, n_cams)))[:, None], k1[:, None], k2[:, None]], 1)
    pts = pts_gt + rng.normal(0, 1.0, (n_pts, 3)) * (pt_sigma * depth)[:, None]
    return BALProblem(cam_idx, pt_idx, np.ascontiguousarray(obs), cams, pts, name)


def make_named(name: str, seed: int = 0) -> BALProblem:
    if name == "long-tracks":
        return make_long_tracks(seed=seed)
    nc, npts, m = SHAPES[name]
    return make_bal(nc, npts, m, seed=seed, name=name)


def make_long_tracks(n_cams: int = 420, n_pts: int = 900, n_obs: int = 6000, tracks=(400, 260, 193, 300),
                     seed: int = 0) -> BALProblem:
    """A BAL problem with LONG TRACKS, as real BAL sets have (landmarks seen by hundreds of cameras): the first point, two
    points in the middle and the last point are observed by ``tracks`` cameras each - more than one tile of the library
    holds (192), so their observations are cut into fragment tiles (csrc/structure.hpp)."""
    base = make_bal(n_cams, n_pts, n_obs, seed=seed, name="long-tracks")
    rng = np.random.Generator(np.random.PCG64(seed + 77))
    if len(tracks) <= 4:
        chosen = [0, n_pts // 3, (2 * n_pts) // 3, n_pts - 1][: len(tracks)]
    else:  # many long tracks: evenly spread over the point range, first and last point included
        chosen = np.unique(np.linspace(0, n_pts - 1, len(tracks)).astype(np.int64)).tolist()
        tracks = tracks[: len(chosen)]
    cam_new, pt_new = [], []
    for p, t in zip(chosen, tracks):
        have = base.cam_idx[base.pt_idx == p]
        others = np.setdiff1d(np.arange(n_cams), have)
        add = rng.choice(others, size=min(max(t - have.size, 0), others.size), replace=False)
        cam_new.append(add.astype(np.int32))
        pt_new.append(np.full(add.size, p, dtype=np.int32))
    cam_new, pt_new = np.concatenate(cam_new), np.concatenate(pt_new)
    cam_idx = np.concatenate([base.cam_idx, cam_new])
    pt_idx = np.concatenate([base.pt_idx, pt_new])
    obs = np.concatenate([base.obs, np.zeros((cam_new.size, 2))])
    # a landmark seen from the whole scene is far away: move the chosen points to a depth of 300 (in front of every camera)
    # and take ALL their observations from the initial estimate's projection plus noise
    pts = base.pts.copy()
    pts[chosen, 2] = -300.0
    sel = np.isin(pt_idx, chosen)
    obs[sel] = project(base.cams, pts, cam_idx[sel], pt_idx[sel]) + rng.normal(0, 2.0, (int(sel.sum()), 2))
    order = np.lexsort((cam_idx, pt_idx))
    return BALProblem(np.ascontiguousarray(cam_idx[order]), np.ascontiguousarray(pt_idx[order]),
                      np.ascontiguousarray(obs[order]), base.cams, pts, "long-tracks")


def schur_fixture() -> BALProblem:
    """The 2-camera / 3-point / 6-observation literals of ``tests/schur.cu:52-78``."""
    cams = np.array([[0.12, -0.08, 0.03, 0.25, -0.10, 0.20, 800.0, 0.01, -0.001],
                     [-0.09, 0.06, -0.04, -0.30, 0.14, -0.22, 820.0, -0.012, 0.0009]])
    # the reference writes the points as float literals (0.1f ...) widened to T
    pts = np.array([[0.1, 0.0, 2.0], [-0.1, 0.05, 2.2], [0.0, -0.05, 1.8]], dtype=np.float32).astype(np.float64)
    cam_idx = np.array([0, 1, 0, 1, 0, 1], dtype=np.int32)
    pt_idx = np.array([0, 0, 1, 1, 2, 2], dtype=np.int32)
    return BALProblem(cam_idx, pt_idx, np.zeros((6, 2)), cams, pts, "schur-fixture")


def precision_matrices(n_obs: int) -> np.ndarray:
    """Deterministic SPD 2x2 precision matrix per observation, [n_obs][2][2], for the weighted / robust test cases.

    Integer arithmetic plus one correctly rounded division per entry, so a C++ restatement of the same
    formula (the driver that generated tests/golden/*__weights.json) produces bit-identical values."""
    i = np.arange(n_obs, dtype=np.int64)
    a = 1.0 + ((i * 7) % 11).astype(np.float64) / 22.0
    b = 0.8 + ((i * 5) % 13).astype(np.float64) / 26.0
    c = (((i * 3) % 7).astype(np.float64) - 3.0) / 20.0
    P = np.empty((n_obs, 2, 2))
    P[:, 0, 0], P[:, 0, 1], P[:, 1, 0], P[:, 1, 1] = a, c, c, b
    return P


def write_gbal(prob: BALProblem, path: str) -> None:
    """Binary container read by the reference driver (int64 header, int32 ids, f64 payload)."""
    with open(path, "wb") as fh:
        fh.write(struct.pack("<qqq", prob.n_cams, prob.n_pts, prob.n_obs))
        fh.write(np.ascontiguousarray(prob.cam_idx, dtype="<i4").tobytes())
        fh.write(np.ascontiguousarray(prob.pt_idx, dtype="<i4").tobytes())
        fh.write(np.ascontiguousarray(prob.obs, dtype="<f8").tobytes())
        fh.write(np.ascontiguousarray(prob.cams, dtype="<f8").tobytes())
        fh.write(np.ascontiguousarray(prob.pts, dtype="<f8").tobytes())


def write_bal_text(prob: BALProblem, path: str) -> None:
    """BAL text format as parsed by ``examples/bal.cu:63-147``."""
    with open(path, "w") as fh:
        fh.write(f"{prob.n_cams} {prob.n_pts} {prob.n_obs}\n")
        for c, p, (u, v) in zip(prob.cam_idx, prob.pt_idx, prob.obs):
            fh.write(f"{c} {p} {u:.17g} {v:.17g}\n")
        for row in prob.cams:
            for v in row:
                fh.write(f"{v:.17g}\n")
        for row in prob.pts:
            for v in row:
                fh.write(f"{v:.17g}\n")


def read_bal_text(path: str) -> BALProblem:
    with open(path) as fh:
        tok = fh.read().split()
    nc, npts, m = int(tok[0]), int(tok[1]), int(tok[2])
    o = np.array(tok[3:3 + 4 * m], dtype=np.float64).reshape(m, 4)
    rest = np.array(tok[3 + 4 * m:], dtype=np.float64)
    cams = rest[: 9 * nc].reshape(nc, 9)
    pts = rest[9 * nc: 9 * nc + 3 * npts].reshape(npts, 3)
    return BALProblem(o[:, 0].astype(np.int32), o[:, 1].astype(np.int32), np.ascontiguousarray(o[:, 2:]), cams, pts, path)


# ---------------------------------------------------------------------------------------------------------------
# pose-graph fixture for the generic factor-graph path (6-dof poses [w, t], between factors 6/6/6, unary priors 6/6)
# ---------------------------------------------------------------------------------------------------------------
@dataclasses.dataclass
class PoseGraph:
    ids: np.ndarray  # int64 [N] global vertex ids (descending: the block order is NOT the insertion order)
    poses: np.ndarray  # float64 [N, 6] initial estimate
    fixed: np.ndarray  # uint8 [N]
    bt_idx: np.ndarray  # int32 [Mb, 2] (i, j) local pose indices
    bt_meas: np.ndarray  # float64 [Mb, 6]
    bt_P: np.ndarray  # float64 [Mb, 6, 6] SPD precision matrices
    bt_active: np.ndarray  # uint8 [Mb] FactorDescriptor::set_active values (the level from which a factor is active)
    pr_idx: np.ndarray  # int32 [Mp]
    pr_meas: np.ndarray  # float64 [Mp, 6]
    huber: float = 0.5  # HuberLoss delta on the between factors


def _aa_to_R(w):
    return _rodrigues(np.asarray(w, dtype=np.float64)[None, :])[0]


def _R_to_aa(R):
    c = (np.trace(R) - 1.0) / 2.0
    th = np.arccos(np.clip(c, -1.0, 1.0))
    if th < 1e-12:
        return np.zeros(3)
    return th / (2.0 * np.sin(th)) * np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])


def pose_graph(n: int = 40, seed: int = 0) -> PoseGraph:
    """Poses on a noisy helix, odometry edges i -> i+1, loop closures i -> i+5 / i+11, priors on every 8th pose.
    Pose 0 is fixed (gauge), pose 3 is fixed as well; the last two between factors are switched off (one from level 1 on, one
    never), so the activity masks matter.  Rational-pattern SPD precision matrices (exact in FP32 as well)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    k = np.arange(n)
    gt = np.zeros((n, 6))
    gt[:, 3] = 4.0 * np.cos(0.35 * k)
    gt[:, 4] = 4.0 * np.sin(0.35 * k)
    gt[:, 5] = 0.1 * k
    gt[:, 0:3] = np.stack([0.05 * np.sin(0.2 * k), 0.04 * np.cos(0.3 * k), 0.35 * k * 0.1 + 0.02], axis=1)
    edges = [(i, i + 1) for i in range(n - 1)] + [(i, i + 5) for i in range(0, n - 5, 3)] + [(i, i + 11) for i in range(0, n - 11, 7)]
    bt_idx = np.array(edges, dtype=np.int32)
    meas = np.zeros((len(edges), 6))
    for e, (i, j) in enumerate(edges):
        Ri, Rj = _aa_to_R(gt[i, :3]), _aa_to_R(gt[j, :3])
        meas[e, :3] = _R_to_aa(Ri.T @ Rj) + rng.normal(0.0, 0.01, 3)
        meas[e, 3:] = Ri.T @ (gt[j, 3:] - gt[i, 3:]) + rng.normal(0.0, 0.02, 3)
    # a few gross outliers so that the Huber branch is taken
    for e in range(7, len(edges), 13):
        meas[e, 3:] += np.array([0.9, -0.7, 0.8])
    P = np.zeros((len(edges), 6, 6))
    for e in range(len(edges)):
        dg = np.array([4.0 + (e * 3) % 5, 4.5 + (e * 5) % 3, 5.0 + (e * 7) % 4, 1.0 + ((e * 2) % 7) / 4.0, 1.25 + ((e * 3) % 5) / 4.0, 1.5 + (e % 3) / 2.0])
        P[e] = np.diag(dg)
        off = ((e * 5) % 9 - 4) / 16.0
        P[e, 0, 3] = P[e, 3, 0] = off
        P[e, 1, 4] = P[e, 4, 1] = -off / 2.0
    active = np.zeros(len(edges), dtype=np.uint8)
    # the switched-off factors are the LAST ones: the reference's chi2 kernel walks the first active_count factors instead of
    # the active index list (ops/chi2.hpp:32-44 as launched by ops/chi2.hpp:46-66), so its cost is only meaningful when
    # the inactive factors form a suffix of the factor array
    active[-2] = 1  # only active from optimisation level 1 on
    active[-1] = 0x7F  # never active below level 127 (FactorDescriptor::set_active keeps only the low 7 bits, factor.hpp:419-430)
    pr_idx = np.arange(0, n, 8, dtype=np.int32)
    pr_meas = gt[pr_idx] + rng.normal(0.0, 0.01, (len(pr_idx), 6))
    init = gt + np.concatenate([rng.normal(0.0, 0.03, (n, 3)), rng.normal(0.0, 0.15, (n, 3))], axis=1)
    fixed = np.zeros(n, dtype=np.uint8)
    fixed[0] = fixed[3] = 1
    init[0] = gt[0]
    return PoseGraph(ids=(1000 - 7 * k).astype(np.int64), poses=init, fixed=fixed, bt_idx=bt_idx, bt_meas=meas, bt_P=P,
                     bt_active=active, pr_idx=pr_idx, pr_meas=pr_meas)


def pose_graph_hard() -> PoseGraph:
    """The same graph started far from the optimum (1.2 rad / 5 units of noise): several LM steps are
    rejected, so backup / revert and the damping schedule are exercised."""
    pg = pose_graph(seed=3)
    rng = np.random.Generator(np.random.PCG64(11))
    n = len(pg.ids) - 1
    pg.poses[1:] += np.concatenate([rng.normal(0, 1.2, (n, 3)), rng.normal(0, 5.0, (n, 3))], axis=1)
    return pg


def write_pose_graph(pg: PoseGraph, path: str) -> None:
    """GPG1 binary (read by the pose-graph driver of the unmodified reference that generates the golden runs): int64 n, mb, mp | f64 huber | i64 ids[n] | f64 poses[6n] | i64 fixed[n] |
    i64 bt_idx[2mb] | f64 bt_meas[6mb] | f64 bt_P[36mb] | i64 bt_active[mb] | i64 pr_idx[mp] | f64 pr_meas[6mp]."""
    with open(path, "wb") as fh:
        fh.write(struct.pack("<qqq", len(pg.ids), len(pg.bt_idx), len(pg.pr_idx)))
        fh.write(struct.pack("<d", pg.huber))
        for a, dt in ((pg.ids, np.int64), (pg.poses, np.float64), (pg.fixed, np.int64), (pg.bt_idx, np.int64), (pg.bt_meas, np.float64),
                      (pg.bt_P, np.float64), (pg.bt_active, np.int64), (pg.pr_idx, np.int64), (pg.pr_meas, np.float64)):
            fh.write(np.ascontiguousarray(a, dtype=dt).tobytes())

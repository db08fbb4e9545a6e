"""Point partitioning for the multi-GPU path (one process per GPU).

The reference is single-GPU; the shard rule is SURVEY.md section 8(e): contiguous ranges of the point-sorted
observation array, balanced by observation count; every rank keeps all cameras.  Only camera-sized vectors
(and a few scalars) are all-reduced by the C library over NCCL.
"""
from __future__ import annotations

import numpy as np

from .synthetic import BALProblem


def point_ranges(pt_idx: np.ndarray, n_pts: int, nranks: int):
    """[(p0, p1)] per rank: contiguous point ranges with near-equal observation counts."""
    counts = np.bincount(pt_idx, minlength=n_pts).astype(np.int64)
    cum = np.concatenate([[0], np.cumsum(counts)])
    total = int(cum[-1])
    bounds = [0]
    for r in range(1, nranks):
        target = total * r / nranks
        p = int(np.searchsorted(cum, target, side="left"))
        p = min(max(p, bounds[-1] + 1), n_pts - (nranks - r))
        bounds.append(p)
    bounds.append(n_pts)
    return [(bounds[r], bounds[r + 1]) for r in range(nranks)]


def partition_by_point(prob: BALProblem, nranks: int, rank: int) -> BALProblem:
    """Sub-problem of one rank: its points (re-indexed from 0), their observations, all cameras."""
    if nranks == 1:
        return prob
    p0, p1 = point_ranges(prob.pt_idx, prob.n_pts, nranks)[rank]
    sel = (prob.pt_idx >= p0) & (prob.pt_idx < p1)
    return BALProblem(prob.cam_idx[sel].copy(), (prob.pt_idx[sel] - p0).astype(np.int32), np.ascontiguousarray(prob.obs[sel]),
                      prob.cams.copy(), np.ascontiguousarray(prob.pts[p0:p1]), f"{prob.name}[rank {rank}/{nranks}]")

"""oracle/oracle_graph.py — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU (numpy, dense, float64) restatement of the reference's GENERIC factor-graph path, used only by tests/ as the checker
of graphite_b200/csrc/graph_generic.cu.  Each method cites the reference lines it follows.  Pinned against the integer
known-answer tests of the reference's tests/factor.cu (tests/test_graph_oracle.py) and against runs of the unmodified
reference on a pose-graph fixture (oracle/ref_pose_driver.cu -> tests/golden/pose-graph*.json).
"""
from __future__ import annotations

import math

import numpy as np


# ------------------------------------------------------------------ forward-mode dual numbers (for the pose factors)
class Dual:
    __slots__ = ("v", "g")

    def __init__(self, v, g):
        self.v, self.g = float(v), g

    @staticmethod
    def lift(x, n):
        return x if isinstance(x, Dual) else Dual(x, np.zeros(n))

    def _o(self, o):
        return o if isinstance(o, Dual) else Dual(o, np.zeros_like(self.g))

    def __add__(self, o):
        o = self._o(o); return Dual(self.v + o.v, self.g + o.g)
    __radd__ = __add__

    def __sub__(self, o):
        o = self._o(o); return Dual(self.v - o.v, self.g - o.g)

    def __rsub__(self, o):
        return self._o(o) - self

    def __mul__(self, o):
        o = self._o(o); return Dual(self.v * o.v, self.g * o.v + self.v * o.g)
    __rmul__ = __mul__

    def __truediv__(self, o):
        o = self._o(o); return Dual(self.v / o.v, (self.g * o.v - self.v * o.g) / (o.v * o.v))

    def __rtruediv__(self, o):
        return self._o(o) / self

    def __neg__(self):
        return Dual(-self.v, -self.g)

    def __gt__(self, o):
        return self.v > (o.v if isinstance(o, Dual) else o)

    def __lt__(self, o):
        return self.v < (o.v if isinstance(o, Dual) else o)


def _sqrt(x):
    return Dual(math.sqrt(x.v), x.g / (2.0 * math.sqrt(x.v))) if isinstance(x, Dual) else math.sqrt(x)


def _sin(x):
    return Dual(math.sin(x.v), x.g * math.cos(x.v)) if isinstance(x, Dual) else math.sin(x)


def _cos(x):
    return Dual(math.cos(x.v), -x.g * math.sin(x.v)) if isinstance(x, Dual) else math.cos(x)


def _acos(x):
    return Dual(math.acos(x.v), -x.g / math.sqrt(1.0 - x.v * x.v)) if isinstance(x, Dual) else math.acos(x)


def pose_rotation(w):
    """tests/user_factor/pose_residual.cuh: pose_rotation (row-major 3x3 as a flat list)."""
    R = [1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0]
    theta = _sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2])
    if theta > 0:
        ax, ay, az = w[0] / theta, w[1] / theta, w[2] / theta
        s, c = _sin(theta), _cos(theta)
        sx, sy, sz = s * ax, s * ay, s * az
        cx, cy, cz = (1.0 - c) * ax, (1.0 - c) * ay, (1.0 - c) * az
        tmp = cx * ay; R[1] = tmp - sz; R[3] = tmp + sz
        tmp = cx * az; R[2] = tmp + sy; R[6] = tmp - sy
        tmp = cy * az; R[5] = tmp - sx; R[7] = tmp + sx
        R[0] = cx * ax + c; R[4] = cy * ay + c; R[8] = cz * az + c
    return R


def between6_residual(xi, xj, z):
    """tests/user_factor/pose_residual.cuh: between6_residual."""
    Ri, Rj, Rz = pose_rotation(xi[:3]), pose_rotation(xj[:3]), pose_rotation([float(z[0]), float(z[1]), float(z[2])])
    M = [Ri[a] * Rj[b] + Ri[3 + a] * Rj[3 + b] + Ri[6 + a] * Rj[6 + b] for a in range(3) for b in range(3)]
    Re = [Rz[a] * M[b] + Rz[3 + a] * M[3 + b] + Rz[6 + a] * M[6 + b] for a in range(3) for b in range(3)]
    c = (Re[0] + Re[4] + Re[8] - 1.0) / 2.0
    k = 0.5
    if c < 1.0:
        theta = _acos(c)
        k = theta / (2.0 * _sin(theta))
    d = [xj[3] - xi[3], xj[4] - xi[4], xj[5] - xi[5]]
    return [k * (Re[7] - Re[5]), k * (Re[2] - Re[6]), k * (Re[3] - Re[1]),
            Ri[0] * d[0] + Ri[3] * d[1] + Ri[6] * d[2] - float(z[3]),
            Ri[1] * d[0] + Ri[4] * d[1] + Ri[7] * d[2] - float(z[4]),
            Ri[2] * d[0] + Ri[5] * d[1] + Ri[8] * d[2] - float(z[5])]


def make_between6(meas):
    def ev(vals, f):
        n = 12
        xi = [Dual(vals[0][k], np.eye(n)[k]) for k in range(6)]
        xj = [Dual(vals[1][k], np.eye(n)[6 + k]) for k in range(6)]
        r = [Dual.lift(x, n) for x in between6_residual(xi, xj, meas[f])]
        J = np.array([x.g for x in r])
        return np.array([x.v for x in r]), [J[:, :6], J[:, 6:]]
    return ev


def make_prior6(meas):
    def ev(vals, f):
        return np.asarray(vals[0], dtype=np.float64) - meas[f], [np.eye(6)]
    return ev


def wrap_pi(a):
    two_pi = 6.283185307179586476925286766559
    return a - two_pi * math.floor((a + 3.14159265358979323846) / two_pi)


def make_se2(meas):
    def ev(vals, f):
        xi, xj, z = vals[0], vals[1], meas[f]
        c, s = math.cos(xi[2]), math.sin(xi[2])
        dx, dy = xj[0] - xi[0], xj[1] - xi[1]
        r = np.array([c * dx + s * dy - z[0], -s * dx + c * dy - z[1], wrap_pi(xj[2] - xi[2] - z[2])])
        Ji = np.array([[-c, -s, -s * dx + c * dy], [s, -c, -c * dx - s * dy], [0, 0, -1.0]])
        Jj = np.array([[c, s, 0], [-s, c, 0], [0, 0, 1.0]])
        return r, [Ji, Jj]
    return ev


def make_linear(A, obs):
    """r = sum_s A_s v_s - obs: the toy factors of tests/factor.cu:8-125."""
    def ev(vals, f):
        r = -np.asarray(obs[f], dtype=np.float64).reshape(-1).copy()
        for s, a in enumerate(A):
            r = r + a @ np.asarray(vals[s], dtype=np.float64)
        return r, [np.array(a, dtype=np.float64) for a in A]
    return ev


# ------------------------------------------------------------------ the graph
class GraphOracle:
    def __init__(self):
        self.V, self.F = [], []
        self.scale_on = True

    def add_vertex_set(self, dimension, global_ids, values, fixed=None, eliminate=False, update=None):
        n = len(global_ids)
        self.V.append(dict(d=dimension, gid=np.asarray(global_ids, dtype=np.int64), x=np.array(values, dtype=np.float64).reshape(n, -1),
                           fixed=np.zeros(n, np.uint8) if fixed is None else np.asarray(fixed, np.uint8), elim=eliminate, update=update))
        return len(self.V) - 1

    def add_factor_set(self, E, vsets, vidx, ev, active=None, P=None, loss=0, delta=0.0):
        vidx = np.asarray(vidx, dtype=np.int64).reshape(-1, len(vsets))
        n = vidx.shape[0]
        self.F.append(dict(E=E, vsets=list(vsets), vidx=vidx, ev=ev, level=np.zeros(n, np.uint8) if active is None else np.asarray(active, np.uint8),
                           P=np.tile(np.eye(E), (n, 1, 1)) if P is None else np.asarray(P, np.float64).reshape(n, E, E), loss=loss, delta=delta))
        return len(self.F) - 1

    # Graph::initialize_optimization (graph.hpp:92-210), Hessian::build_structure (hessian.hpp:48-85, 257-288)
    def initialize(self, level=0):
        used = [np.zeros(len(v["gid"]), bool) for v in self.V]
        self.rows = 0
        for f in self.F:
            a = f["level"]
            f["active"] = ((a & 0x7F) <= level) & ((a & 0x80) == 0)  # active.hpp:11-16
            for i in np.nonzero(f["active"])[0]:
                for s, vs in enumerate(f["vsets"]):
                    used[vs][f["vidx"][i, s]] = True
            f["roff"] = self.rows
            self.rows += f["E"] * f["vidx"].shape[0]
        order = sorted(((bool(v["elim"]), int(g), si, k) for si, v in enumerate(self.V) for k, g in enumerate(v["gid"])))
        self.dimH, self.nblocks, self.block_dim = 0, 0, []
        for v in self.V:
            v["hoff"] = -np.ones(len(v["gid"]), np.int64)
            v["block"] = -np.ones(len(v["gid"]), np.int64)
        for _, _, si, k in order:
            v = self.V[si]
            if v["fixed"][k] or not used[si][k]:
                continue
            v["hoff"][k], v["block"][k] = self.dimH, self.nblocks
            self.block_dim.append(v["d"])
            self.dimH += v["d"]
            self.nblocks += 1
        coords = set()
        for f in self.F:
            for i in np.nonzero(f["active"])[0]:
                for a in range(len(f["vsets"])):
                    for b in range(a, len(f["vsets"])):
                        ba = self.V[f["vsets"][a]]["block"][f["vidx"][i, a]]
                        bb = self.V[f["vsets"][b]]["block"][f["vidx"][i, b]]
                        if ba >= 0 and bb >= 0:
                            coords.add((max(ba, bb), min(ba, bb)))  # (col, row)
        coords = sorted(coords)
        colptr = np.zeros(self.nblocks + 1, np.int64)
        rowidx, offsets, nv = [], [], 0
        for c, r in coords:
            colptr[c + 1] += 1
            rowidx.append(r)
            offsets.append(nv)
            nv += self.block_dim[r] * self.block_dim[c]
        self.colptr, self.rowidx, self.offsets, self.nvalues = np.cumsum(colptr), np.array(rowidx, np.int64), np.array(offsets, np.int64), nv
        self.coords = coords
        return self.dimH

    def _loss(self, f, c):
        if f["loss"] == 1 and c > f["delta"] ** 2:  # loss.hpp:36-50
            return 2.0 * math.sqrt(c) * f["delta"] - f["delta"] ** 2, f["delta"] / math.sqrt(c)
        return c, 1.0

    def _evaluate(self):
        """residuals, unscaled dense Jacobian, block weights dL*P, chi2 (ops/error.hpp, ops/linearize.hpp:8-138, ops/chi2.hpp:9-44)."""
        J = np.zeros((self.rows, self.dimH))
        r = np.zeros(self.rows)
        W = np.zeros((self.rows, self.rows))
        chi2 = 0.0
        for f in self.F:
            E = f["E"]
            f["chi2"] = np.zeros(f["vidx"].shape[0])
            f["dL"] = np.zeros(f["vidx"].shape[0])
            for i in np.nonzero(f["active"])[0]:
                vals = [self.V[vs]["x"][f["vidx"][i, s]] for s, vs in enumerate(f["vsets"])]
                ri, Js = f["ev"](vals, i)
                row = f["roff"] + E * i
                r[row:row + E] = ri
                c, dl = self._loss(f, float(ri @ f["P"][i] @ ri))
                f["chi2"][i], f["dL"][i] = c, dl
                chi2 += c
                W[row:row + E, row:row + E] = dl * f["P"][i]
                for s, vs in enumerate(f["vsets"]):
                    ho = self.V[vs]["hoff"][f["vidx"][i, s]]
                    if ho >= 0:
                        J[row:row + E, ho:ho + self.V[vs]["d"]] += Js[s]
        return r, J, W, chi2

    # Graph::linearize (graph.hpp:236-290)
    def linearize(self):
        self.r, J, self.W, self.chi2 = self._evaluate()
        d = np.einsum("ij,ik,kj->j", J, self.W, J)
        self.scales = 1.0 / (np.finfo(np.float64).eps + np.sqrt(d)) if self.scale_on else np.ones(self.dimH)
        self.J = J * self.scales
        self.b = -self.J.T @ (self.W @ self.r)
        self.H = self.J.T @ self.W @ self.J
        return self.chi2

    def cost(self):
        chi2 = 0.0
        for f in self.F:
            for i in np.nonzero(f["active"])[0]:
                vals = [self.V[vs]["x"][f["vidx"][i, s]] for s, vs in enumerate(f["vsets"])]
                ri, _ = f["ev"](vals, i)
                chi2 += self._loss(f, float(ri @ f["P"][i] @ ri))[0]
        return chi2

    def hessian_values(self):
        """upper block-CSC values, blocks column-major (hessian.hpp:59-84, ops/hessian.hpp:58-76)."""
        out = np.zeros(self.nvalues)
        starts = np.concatenate([[0], np.cumsum(self.block_dim)])
        for (c, r), off in zip(self.coords, self.offsets):
            blk = self.H[starts[r]:starts[r + 1], starts[c]:starts[c + 1]]
            out[off:off + blk.size] = blk.T.reshape(-1)
        return out

    def jv(self, x):
        return self.J @ x

    def jtpv(self, v):
        return self.J.T @ (self.W @ v)

    def block_inverse(self, mu, identity):
        """BlockJacobiPreconditioner (preconditioner/block_jacobi.hpp:79-168, ops/hessian.hpp:80-109)."""
        M = np.zeros((self.dimH, self.dimH))
        o = 0
        for d in self.block_dim:
            blk = self.H[o:o + d, o:o + d].copy()
            dg = np.diag(blk).copy()
            np.fill_diagonal(blk, dg + mu if identity else dg + mu * np.clip(dg, 1e-6, 1e32))
            M[o:o + d, o:o + d] = np.linalg.inv(blk)
            o += d
        return M

    # PCGSolver::solve (solver/pcg.hpp:61-232)
    def solve(self, mu, identity=False, max_iter=10, tol=1.0, ratio=5.0):
        Minv = self.block_inverse(mu, identity)
        diag = np.clip(np.diag(self.H), 1e-6, 1e32)
        x = np.zeros(self.dimH)
        r = self.b.copy()
        z = Minv @ (r / np.sqrt(r @ r))
        p = z.copy()
        rz, rz0, k = r @ z, np.inf, 0
        for k in range(1, max_iter + 1):
            if rz == 0:
                k -= 1
                break
            v2 = self.jtpv(self.jv(p)) + (mu * p if identity else mu * diag * p)
            alpha = rz / (p @ v2)
            xb = x.copy()
            x = x + alpha * p
            r = r - alpha * v2
            z = Minv @ (r / np.sqrt(r @ r))
            rz_new = r @ z
            if abs(rz_new) > ratio * rz0 or np.isnan(rz_new):
                x = xb
                break
            rz0 = min(rz0, abs(rz_new))
            beta = rz_new / rz
            rz = rz_new
            p = z + beta * p
            if abs(rz_new) < tol:
                break
        return x, k

    def _apply(self, x):
        for v in self.V:
            on = v["hoff"] >= 0
            for k in np.nonzero(on)[0]:
                ho, d = v["hoff"][k], v["d"]
                delta = x[ho:ho + d] * self.scales[ho:ho + d]  # ops/update.hpp:23-30
                if v["update"] is not None:
                    v["x"][k] = v["update"](v["x"][k], delta)
                else:
                    v["x"][k, :d] += delta

    # optimizer::levenberg_marquardt (levenberg_marquardt.hpp:109-242, compute_rho :19-47)
    def lm(self, iterations=50, initial_damping=1e-4, identity=False, pcg_iterations=10, pcg_tolerance=1.0, rejection_ratio=5.0):
        mu, nu = initial_damping, 2.0
        chi2 = self.linearize()
        table = []
        for _ in range(iterations):
            x, k = self.solve(mu, identity, pcg_iterations, pcg_tolerance, rejection_ratio)
            backup = [v["x"].copy() for v in self.V]
            self._apply(x)
            new_chi2 = self.cost()
            rho = (chi2 - new_chi2) / (float(np.sum(x * (mu * x + self.b))) + 1e-3)
            if np.isfinite(new_chi2) and rho > 0:
                alpha = min(max(1.0 - (2.0 * rho - 1.0) ** 3, 1.0 / 3.0), 2.0 / 3.0)
                mu *= alpha
                nu = 2.0
                self.linearize()
            else:
                for v, bk in zip(self.V, backup):
                    v["x"] = bk
                mu *= nu
                nu *= 2.0
                new_chi2 = chi2
            table.append([chi2, new_chi2, mu, k])
            chi2 = new_chi2
            if not np.isfinite(mu) or rho == 0:
                break
        return np.array(table)

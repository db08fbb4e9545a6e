// oracle/oracle_bal.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// CPU restatement (plain C++17 + OpenMP, no Eigen) of the reference's
// Levenberg-Marquardt inner loop for bundle adjustment.  It exists to CHECK the
// CUDA product path and to provide the CPU baseline timing; nothing under
// graphite_b200/ links, loads or calls it.
//
// Parity status: PINNED.  tests/test_oracle_golden.py checks this file against
// golden outputs of the unmodified reference run on a B200 (oracle/_ref/ref_bal,
// fixtures in tests/golden/, generator oracle/make_golden.py) and against the
// reference's own test literals (tests/schur.cu:52-78 fixture).
//
// Reference map (paths relative to /root/reference):
//   residual               examples/reprojection_error.cuh:61-99
//   analytic Jacobian      examples/projection_jacobians.cuh:2-322 (same function, derived by hand
//                          here: chain rule through Rodrigues; theta==0 gives zero d/dw as the
//                          reference's else-branch does, :200-236)
//   chi2                   include/graphite/ops/chi2.hpp:9-44, factor.hpp:551-557 (no 1/2 factor)
//   linearize/scales/b     include/graphite/graph.hpp:236-290, ops/linearize.hpp:140-180,238-303,
//                          ops/hessian.hpp:418-474
//   Hessian blocks/layout  include/graphite/ops/hessian.hpp:9-78, hessian.hpp:257-288, csc_utils.hpp:16-50
//   damping                include/graphite/hessian.hpp:136-176
//   Schur, b_S, back-subst include/graphite/schur.hpp:227-302, ops/schur.hpp:154-188, tests/schur_cpu_ref.cpp:8-51
//   block-Jacobi           include/graphite/preconditioner/block_jacobi_schur.hpp:114-178
//   PCG                    include/graphite/solver/pcg_schur.hpp:79-168
//   full-system PCG        include/graphite/solver/pcg.hpp:61-232, preconditioner/block_jacobi.hpp:79-186
//   direct solve           include/graphite/solver/eigen_schur.hpp:52-108, src/eigen_solver.cpp:10-29 (as dense LDL^T)
//   update/backup/revert   include/graphite/ops/update.hpp:9-31, graph.hpp:292-318
//   rho and LM control     include/graphite/optimizer/levenberg_marquardt.hpp:19-47,109-242
#include "oracle_bal.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

using clk = std::chrono::steady_clock;
static double secs(clk::time_point a) { return std::chrono::duration<double>(clk::now() - a).count(); }

struct Base {
  virtual ~Base() {}
  virtual void set_threads(int) = 0;
  virtual void get_params(double *, double *) = 0;
  virtual void set_params(const double *, const double *) = 0;
  virtual void set_robust(int, double, const double *) = 0;
  virtual double residuals(double *) = 0;
  virtual void jacobians(double *, double *) = 0;
  virtual double linearize(double *, double *) = 0;
  virtual void hessian_structure(int64_t *, int64_t *, int64_t *) = 0;
  virtual int64_t hessian_num_values() = 0;
  virtual void hessian_values(double *) = 0;
  virtual void schur(double, int, double *, double *) = 0;
  virtual int64_t schur_nnz_blocks() = 0;
  virtual int64_t solve(const orc_lm_options *, double, double *) = 0;
  virtual int64_t lm(const orc_lm_options *, double *) = 0;
  virtual void lm_begin(const orc_lm_options *) = 0;
  virtual int lm_step(const orc_lm_options *, double *) = 0;
  double tim[6] = {0, 0, 0, 0, 0, 0};
};

// ---- per-observation math -------------------------------------------------------------
// examples/reprojection_error.cuh:61-99 (Eigen::AngleAxis::toRotationMatrix written out).
template <typename T> static inline void rotation(const T *w, T *R /*row-major*/, T &theta) {
  theta = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  R[0] = R[4] = R[8] = T(1);
  R[1] = R[2] = R[3] = R[5] = R[6] = R[7] = T(0);
  if (theta > T(0)) {
    const T ax = w[0] / theta, ay = w[1] / theta, az = w[2] / theta;
    const T s = std::sin(theta), c = std::cos(theta);
    const T sx = s * ax, sy = s * ay, sz = s * az;
    const T cx = (T(1) - c) * ax, cy = (T(1) - c) * ay, cz = (T(1) - c) * az;
    T tmp;
    tmp = cx * ay; R[1] = tmp - sz; R[3] = tmp + sz;
    tmp = cx * az; R[2] = tmp + sy; R[6] = tmp - sy;
    tmp = cy * az; R[5] = tmp - sx; R[7] = tmp + sx;
    R[0] = cx * ax + c; R[4] = cy * ay + c; R[8] = cz * az + c;
  }
}

template <typename T> static inline void residual(const T *cam, const T *X, const T *obs, T *r) {
  T R[9], theta;
  rotation(cam, R, theta);
  const T Px = R[0] * X[0] + R[1] * X[1] + R[2] * X[2] + cam[3];
  const T Py = R[3] * X[0] + R[4] * X[1] + R[5] * X[2] + cam[4];
  const T Pz = R[6] * X[0] + R[7] * X[1] + R[8] * X[2] + cam[5];
  const T px = -Px / Pz, py = -Py / Pz;
  const T r2 = px * px + py * py;
  const T rd = T(1) + cam[7] * r2 + cam[8] * r2 * r2;
  r[0] = cam[6] * rd * px - obs[0];
  r[1] = cam[6] * rd * py - obs[1];
}

// d r / d(cam, point), column-major 2x9 and 2x3 (ops/linearize.hpp:36-38).  Same function as
// examples/projection_jacobians.cuh; derived here as  J_t = G,  J_w = G * d(RX)/dw,  J_X = G * R
// with G = dr/dP (2x3).
template <typename T> static inline void jacobian(const T *cam, const T *X, T *Jc, T *Jp) {
  T R[9], theta;
  rotation(cam, R, theta);
  const T *w = cam;
  const T Px = R[0] * X[0] + R[1] * X[1] + R[2] * X[2] + cam[3];
  const T Py = R[3] * X[0] + R[4] * X[1] + R[5] * X[2] + cam[4];
  const T Pz = R[6] * X[0] + R[7] * X[1] + R[8] * X[2] + cam[5];
  const T iz = T(1) / Pz;
  const T px = -Px * iz, py = -Py * iz;
  const T r2 = px * px + py * py;
  const T f = cam[6], k1 = cam[7], k2 = cam[8];
  const T d = T(1) + k1 * r2 + k2 * r2 * r2;
  const T e = T(2) * k1 + T(4) * k2 * r2;
  // G = f (d I + e p p^T) [[-iz,0,-px iz],[0,-iz,-py iz]]
  T G[6]; // row-major 2x3
  G[0] = -f * iz * (d + e * px * px);
  G[1] = -f * iz * (e * px * py);
  G[2] = -f * iz * px * (d + e * r2);
  G[3] = -f * iz * (e * px * py);
  G[4] = -f * iz * (d + e * py * py);
  G[5] = -f * iz * py * (d + e * r2);
  // D = d(R X)/dw (3x3, row-major), zero when theta == 0 (reference else-branch)
  T D[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (theta > T(0)) {
    const T t2 = theta * theta;
    const T c = std::cos(theta), s = std::sin(theta);
    const T A = s / theta, B = (T(1) - c) / t2;
    T Ap, Bp;
    if (t2 < T(1e-4)) { // series: avoids the cancellation in (c - A) and (A - 2B)
      Ap = T(-1.0 / 3.0) + t2 * (T(1.0 / 30.0) - t2 * T(1.0 / 840.0));
      Bp = T(-1.0 / 12.0) + t2 * (T(1.0 / 180.0) - t2 * T(1.0 / 6720.0));
    } else {
      Ap = (c - A) / t2;
      Bp = (A - T(2) * B) / t2;
    }
    const T wx[3] = {w[1] * X[2] - w[2] * X[1], w[2] * X[0] - w[0] * X[2], w[0] * X[1] - w[1] * X[0]};
    const T wX = w[0] * X[0] + w[1] * X[1] + w[2] * X[2];
    T u[3];
    for (int i = 0; i < 3; i++) u[i] = -A * X[i] + Ap * wx[i] + Bp * wX * w[i];
    // -[X]x
    const T mXx[9] = {0, X[2], -X[1], -X[2], 0, X[0], X[1], -X[0], 0};
    for (int i = 0; i < 3; i++)
      for (int k = 0; k < 3; k++)
        D[i * 3 + k] = u[i] * w[k] + A * mXx[i * 3 + k] + B * ((i == k ? wX : T(0)) + w[i] * X[k]);
  }
  for (int row = 0; row < 2; row++) {
    const T *g = G + 3 * row;
    for (int k = 0; k < 3; k++) {
      Jc[row + 2 * k] = g[0] * D[k] + g[1] * D[3 + k] + g[2] * D[6 + k];
      Jc[row + 2 * (3 + k)] = g[k];
      Jp[row + 2 * k] = g[0] * R[k] + g[1] * R[3 + k] + g[2] * R[6 + k];
    }
  }
  Jc[0 + 12] = d * px; Jc[1 + 12] = d * py;
  Jc[0 + 14] = f * r2 * px; Jc[1 + 14] = f * r2 * py;
  Jc[0 + 16] = f * r2 * r2 * px; Jc[1 + 16] = f * r2 * r2 * py;
}

// Dense symmetric positive-definite inverse by Gauss-Jordan with partial pivoting (n = 3 or 9);
// stands in for cublas<t>matinvBatched (schur.hpp:1100-1110, block_jacobi_schur.hpp:139-147).
template <typename T, int N> static inline void invert(const T *A /*col-major*/, T *Ai) {
  T M[N][2 * N];
  for (int i = 0; i < N; i++)
    for (int j = 0; j < N; j++) {
      M[i][j] = A[i + j * N];
      M[i][N + j] = (i == j) ? T(1) : T(0);
    }
  for (int c = 0; c < N; c++) {
    int piv = c;
    for (int i = c + 1; i < N; i++)
      if (std::fabs(M[i][c]) > std::fabs(M[piv][c])) piv = i;
    if (piv != c)
      for (int j = 0; j < 2 * N; j++) std::swap(M[c][j], M[piv][j]);
    const T ip = T(1) / M[c][c];
    for (int j = 0; j < 2 * N; j++) M[c][j] *= ip;
    for (int i = 0; i < N; i++)
      if (i != c) {
        const T fct = M[i][c];
        if (fct != T(0))
          for (int j = 0; j < 2 * N; j++) M[i][j] -= fct * M[c][j];
      }
  }
  for (int i = 0; i < N; i++)
    for (int j = 0; j < N; j++) Ai[i + j * N] = M[i][N + j];
}

template <typename T> struct Impl : Base {
  int64_t nc, np, m, dimc, dimH;
  std::vector<int32_t> ci, pi;
  std::vector<int64_t> pptr; // CSR by point over the (point,camera)-sorted observations
  std::vector<T> obs, cams, pts, cams_bak, pts_bak;
  std::vector<T> r, Jc, Jp; // J scaled after linearize (S == T)
  // per-factor precision matrix P (row-major 2x2, identity by default: factor.hpp:397-405) and loss
  // (loss.hpp:15-51: 0 = DefaultLoss, 1 = HuberLoss(delta)); dLv = loss'(r^T P r) of the last error evaluation
  std::vector<T> Pm, dLv;
  int loss_kind = 0;
  T loss_delta = T(0);
  std::vector<T> scales, b;
  // Hessian blocks in the scaled space (undamped): B [nc][81], E [m][27], C [np][9], col-major blocks
  std::vector<T> Bk, Ek, Ck, Bd, Cd, Cinv;
  // Schur: upper block CSR
  std::vector<int64_t> srow;
  std::vector<int32_t> scol;
  std::vector<T> Sv, bS, Minv;
  int threads = 0;
  bool have_structure = false;

  Impl(int64_t nc_, int64_t np_, int64_t m_, const int32_t *c, const int32_t *p, const double *o, const double *cm,
       const double *pt)
      : nc(nc_), np(np_), m(m_), dimc(9 * nc_), dimH(9 * nc_ + 3 * np_), ci(c, c + m_), pi(p, p + m_) {
    obs.resize(2 * m); cams.resize(9 * nc); pts.resize(3 * np);
    for (int64_t i = 0; i < 2 * m; i++) obs[i] = (T)o[i];
    for (int64_t i = 0; i < 9 * nc; i++) cams[i] = (T)cm[i];
    for (int64_t i = 0; i < 3 * np; i++) pts[i] = (T)pt[i];
    pptr.assign(np + 1, 0);
    for (int64_t i = 0; i < m; i++) pptr[pi[i] + 1]++;
    for (int64_t i = 0; i < np; i++) pptr[i + 1] += pptr[i];
    r.resize(2 * m); Jc.resize(18 * m); Jp.resize(6 * m);
    Pm.assign(4 * m, T(0)); dLv.assign(m, T(1));
    for (int64_t f = 0; f < m; f++) Pm[4 * f] = Pm[4 * f + 3] = T(1);
    scales.resize(dimH); b.resize(dimH);
#ifdef _OPENMP
    threads = omp_get_max_threads();
#else
    threads = 1;
#endif
  }
  void set_threads(int t) override {
#ifdef _OPENMP
    threads = t > 0 ? t : omp_get_max_threads();
#else
    (void)t; threads = 1;
#endif
  }
  void get_params(double *c, double *p) override {
    for (int64_t i = 0; i < 9 * nc; i++) c[i] = cams[i];
    for (int64_t i = 0; i < 3 * np; i++) p[i] = pts[i];
  }
  void set_params(const double *c, const double *p) override {
    for (int64_t i = 0; i < 9 * nc; i++) cams[i] = (T)c[i];
    for (int64_t i = 0; i < 3 * np; i++) pts[i] = (T)p[i];
  }

  void set_robust(int kind, double delta, const double *P) override {
    loss_kind = kind;
    loss_delta = (T)delta;
    for (int64_t f = 0; f < m; f++)
      for (int i = 0; i < 4; i++) Pm[4 * f + i] = P ? (T)P[4 * f + i] : (i == 0 || i == 3 ? T(1) : T(0));
  }
  // a^T P b for 2-vectors, in the operation order of ops/hessian.hpp:58-72 / ops/product.hpp:270-281
  static inline T pw(const T *a, const T *b, const T *P) {
    T v = 0;
    for (int i = 0; i < 2; i++) {
      T pj = 0;
      for (int j = 0; j < 2; j++) pj += P[i * 2 + j] * b[j];
      v += a[i] * pj;
    }
    return v;
  }
  // loss.hpp:20-50
  T loss_value(T x) const {
    if (loss_kind == 0 || x <= loss_delta * loss_delta) return x;
    return 2 * std::sqrt(x) * loss_delta - loss_delta * loss_delta;
  }
  T loss_derivative(T x) const {
    if (loss_kind == 0 || x <= loss_delta * loss_delta) return T(1);
    return loss_delta / std::sqrt(x);
  }

  // ops/error.hpp:250-323 then ops/chi2.hpp:9-44: chi2_f = loss(r^T P r), dL_f = loss'(r^T P r).
  T compute_error_chi2() {
    double total = 0; // thrust::reduce over T chi2_vec; accumulate wide here, cast below
    T tot_T = 0;
#pragma omp parallel for num_threads(threads) schedule(static) reduction(+ : total)
    for (int64_t f = 0; f < m; f++) {
      residual(&cams[9 * ci[f]], &pts[3 * pi[f]], &obs[2 * f], &r[2 * f]);
      const T raw = pw(&r[2 * f], &r[2 * f], &Pm[4 * f]);
      dLv[f] = loss_derivative(raw);
      total += (double)loss_value(raw);
    }
    tot_T = (T)total;
    return tot_T;
  }
  double residuals(double *out) override {
    T c = compute_error_chi2();
    if (out)
      for (int64_t i = 0; i < 2 * m; i++) out[i] = r[i];
    return c;
  }
  void raw_jacobians() {
#pragma omp parallel for num_threads(threads) schedule(static)
    for (int64_t f = 0; f < m; f++) jacobian(&cams[9 * ci[f]], &pts[3 * pi[f]], &Jc[18 * f], &Jp[6 * f]);
  }
  void jacobians(double *oc, double *op) override {
    raw_jacobians();
    for (int64_t i = 0; i < 18 * m; i++) oc[i] = Jc[i];
    for (int64_t i = 0; i < 6 * m; i++) op[i] = Jp[i];
  }

  // x2 = dL P r (ops/linearize.hpp:277-291)
  void weighted_residual(int64_t f, T *x2) const {
    for (int i = 0; i < 2; i++) {
      x2[i] = 0;
      for (int j = 0; j < 2; j++) x2[i] += dLv[f] * Pm[4 * f + 2 * i + j] * r[2 * f + j];
    }
  }

  // graph.hpp:236-290
  T do_linearize() {
    auto t0 = clk::now();
    T chi2 = compute_error_chi2();
    raw_jacobians();
    // scalar diagonal, ops/hessian.hpp:418-474 (dL = 1, P = I)
    std::vector<T> diag(dimH, T(0));
    // camera part: reduce per camera deterministically (reference: atomics)
    std::vector<std::vector<T>> tl(threads, std::vector<T>(dimc, T(0)));
#pragma omp parallel num_threads(threads)
    {
#ifdef _OPENMP
      const int tid = omp_get_thread_num();
#else
      const int tid = 0;
#endif
      T *d = tl[tid].data();
#pragma omp for schedule(static)
      for (int64_t f = 0; f < m; f++) {
        const T *J = &Jc[18 * f];
        T *dd = d + 9 * ci[f];
        for (int k = 0; k < 9; k++) dd[k] += pw(&J[2 * k], &J[2 * k], &Pm[4 * f]) * dLv[f];
      }
    }
    for (int t = 0; t < threads; t++)
      for (int64_t i = 0; i < dimc; i++) diag[i] += tl[t][i];
#pragma omp parallel for num_threads(threads) schedule(static)
    for (int64_t p = 0; p < np; p++) {
      T acc[3] = {0, 0, 0};
      for (int64_t f = pptr[p]; f < pptr[p + 1]; f++) {
        const T *J = &Jp[6 * f];
        for (int k = 0; k < 3; k++) acc[k] += pw(&J[2 * k], &J[2 * k], &Pm[4 * f]) * dLv[f];
      }
      for (int k = 0; k < 3; k++) diag[dimc + 3 * p + k] = acc[k];
    }
    // graph.hpp:262-270
    for (int64_t i = 0; i < dimH; i++) {
      const double denom = std::numeric_limits<double>::epsilon() + std::sqrt((double)diag[i]);
      scales[i] = (T)(1.0 / denom);
    }
    // ops/linearize.hpp:140-180
#pragma omp parallel for num_threads(threads) schedule(static)
    for (int64_t f = 0; f < m; f++) {
      const T *sc = &scales[9 * ci[f]], *sp = &scales[dimc + 3 * pi[f]];
      for (int k = 0; k < 9; k++) { Jc[18 * f + 2 * k] *= sc[k]; Jc[18 * f + 2 * k + 1] *= sc[k]; }
      for (int k = 0; k < 3; k++) { Jp[6 * f + 2 * k] *= sp[k]; Jp[6 * f + 2 * k + 1] *= sp[k]; }
    }
    // b = -J~^T r, ops/linearize.hpp:238-303
    std::fill(b.begin(), b.end(), T(0));
    for (auto &v : tl) std::fill(v.begin(), v.end(), T(0));
#pragma omp parallel num_threads(threads)
    {
#ifdef _OPENMP
      const int tid = omp_get_thread_num();
#else
      const int tid = 0;
#endif
      T *d = tl[tid].data();
#pragma omp for schedule(static)
      for (int64_t f = 0; f < m; f++) {
        const T *J = &Jc[18 * f];
        T *dd = d + 9 * ci[f];
        T x2[2];
        weighted_residual(f, x2);
        for (int k = 0; k < 9; k++) dd[k] -= J[2 * k] * x2[0] + J[2 * k + 1] * x2[1];
      }
    }
    for (int t = 0; t < threads; t++)
      for (int64_t i = 0; i < dimc; i++) b[i] += tl[t][i];
#pragma omp parallel for num_threads(threads) schedule(static)
    for (int64_t p = 0; p < np; p++) {
      T acc[3] = {0, 0, 0};
      for (int64_t f = pptr[p]; f < pptr[p + 1]; f++) {
        const T *J = &Jp[6 * f];
        T x2[2];
        weighted_residual(f, x2);
        for (int k = 0; k < 3; k++) acc[k] -= J[2 * k] * x2[0] + J[2 * k + 1] * x2[1];
      }
      for (int k = 0; k < 3; k++) b[dimc + 3 * p + k] = acc[k];
    }
    tim[0] += secs(t0);
    return chi2;
  }
  double linearize(double *sc, double *bb) override {
    T c = do_linearize();
    if (sc) for (int64_t i = 0; i < dimH; i++) sc[i] = scales[i];
    if (bb) for (int64_t i = 0; i < dimH; i++) bb[i] = b[i];
    return c;
  }

  // hessian.hpp:257-288: coordinates sorted by (col,row); offsets = running sum in that order.
  void hessian_structure(int64_t *colptr, int64_t *rowidx, int64_t *offsets) override {
    int64_t k = 0, off = 0;
    for (int64_t c = 0; c < nc; c++) {
      colptr[c] = k; rowidx[k] = c; offsets[k] = off; off += 81; k++;
    }
    for (int64_t p = 0; p < np; p++) {
      colptr[nc + p] = k;
      for (int64_t f = pptr[p]; f < pptr[p + 1]; f++) { rowidx[k] = ci[f]; offsets[k] = off; off += 27; k++; }
      rowidx[k] = nc + p; offsets[k] = off; off += 9; k++;
    }
    colptr[nc + np] = k;
  }
  int64_t hessian_num_values() override { return 81 * nc + 27 * m + 9 * np; }

  // ops/hessian.hpp:9-78: H_ij += J_i^T J_j (dL = 1, P = I), blocks column-major dim_i x dim_j.
  void build_hessian() {
    auto t0 = clk::now();
    Bk.assign(81 * nc, T(0)); Ek.resize(27 * m); Ck.assign(9 * np, T(0));
    std::vector<std::vector<T>> tl(threads);
#pragma omp parallel num_threads(threads)
    {
#ifdef _OPENMP
      const int tid = omp_get_thread_num();
#else
      const int tid = 0;
#endif
      tl[tid].assign(81 * nc, T(0));
      T *Bt = tl[tid].data();
#pragma omp for schedule(static)
      for (int64_t f = 0; f < m; f++) {
        const T *J = &Jc[18 * f];
        T *Bb = Bt + 81 * ci[f];
        for (int j = 0; j < 9; j++)
          for (int i = 0; i < 9; i++) Bb[i + 9 * j] += pw(&J[2 * i], &J[2 * j], &Pm[4 * f]) * dLv[f];
      }
    }
    for (int t = 0; t < threads; t++)
      for (int64_t i = 0; i < 81 * nc; i++) Bk[i] += tl[t][i];
#pragma omp parallel for num_threads(threads) schedule(static)
    for (int64_t p = 0; p < np; p++) {
      T *Cb = &Ck[9 * p];
      for (int64_t f = pptr[p]; f < pptr[p + 1]; f++) {
        const T *J = &Jc[18 * f], *Q = &Jp[6 * f];
        T *Eb = &Ek[27 * f];
        for (int j = 0; j < 3; j++)
          for (int i = 0; i < 9; i++) Eb[i + 9 * j] = pw(&J[2 * i], &Q[2 * j], &Pm[4 * f]) * dLv[f];
        for (int j = 0; j < 3; j++)
          for (int i = 0; i < 3; i++) Cb[i + 3 * j] += pw(&Q[2 * i], &Q[2 * j], &Pm[4 * f]) * dLv[f];
      }
    }
    // hessian.hpp:102-134 backup_diagonal
    Bd.resize(dimc); Cd.resize(3 * np);
    for (int64_t c = 0; c < nc; c++) for (int k = 0; k < 9; k++) Bd[9 * c + k] = Bk[81 * c + 10 * k];
    for (int64_t p = 0; p < np; p++) for (int k = 0; k < 3; k++) Cd[3 * p + k] = Ck[9 * p + 4 * k];
    tim[1] += secs(t0);
  }
  void hessian_values(double *v) override {
    build_hessian();
    int64_t off = 0;
    for (int64_t i = 0; i < 81 * nc; i++) v[off++] = Bk[i];
    for (int64_t p = 0; p < np; p++) {
      for (int64_t f = pptr[p]; f < pptr[p + 1]; f++) for (int i = 0; i < 27; i++) v[off++] = Ek[27 * f + i];
      for (int i = 0; i < 9; i++) v[off++] = Ck[9 * p + i];
    }
  }

  // hessian.hpp:146-175
  static T damp(T d, T mu, bool ident) {
    if (ident) return (T)((double)d + (double)mu);
    return (T)((double)d + mu * std::clamp((double)d, 1.0e-6, 1.0e32));
  }

  // schur.hpp:397-476: S sparsity = Hpp U {(i,j): i<=j co-observe a point}
  void build_schur_structure() {
    if (have_structure) return;
    std::vector<std::vector<int32_t>> rows(nc);
    for (int64_t c = 0; c < nc; c++) rows[c].push_back((int32_t)c);
    for (int64_t p = 0; p < np; p++)
      for (int64_t a = pptr[p]; a < pptr[p + 1]; a++)
        for (int64_t bb = a; bb < pptr[p + 1]; bb++) rows[ci[a]].push_back(ci[bb]);
    srow.assign(nc + 1, 0);
    for (int64_t c = 0; c < nc; c++) {
      auto &v = rows[c];
      std::sort(v.begin(), v.end());
      v.erase(std::unique(v.begin(), v.end()), v.end());
      srow[c + 1] = srow[c] + (int64_t)v.size();
    }
    scol.resize(srow[nc]);
    for (int64_t c = 0; c < nc; c++) std::copy(rows[c].begin(), rows[c].end(), scol.begin() + srow[c]);
    have_structure = true;
  }
  int64_t schur_nnz_blocks() override { build_schur_structure(); return srow[nc]; }
  int64_t sfind(int64_t i, int32_t j) const {
    const int32_t *b0 = &scol[srow[i]], *e0 = &scol[srow[i + 1]];
    return srow[i] + (std::lower_bound(b0, e0, j) - b0);
  }

  // schur.hpp:227-235 on the damped H: S = Hpp - Hpl Hll^-1 Hpl^T (upper blocks), b_S = b_p - Hpl Hll^-1 b_l
  void build_schur(T mu, bool ident) {
    auto t0 = clk::now();
    build_schur_structure();
    Sv.assign(81 * srow[nc], T(0));
    bS.assign(dimc, T(0));
    Cinv.resize(9 * np);
    // execute_Hpp_copy (:587-614) with damped diagonal
    for (int64_t c = 0; c < nc; c++) {
      T *d = &Sv[81 * sfind(c, (int32_t)c)];
      for (int i = 0; i < 81; i++) d[i] = Bk[81 * c + i];
      for (int k = 0; k < 9; k++) d[10 * k] = damp(Bd[9 * c + k], mu, ident);
    }
    // execute_block_diagonal_inversion (:1067-1114)
#pragma omp parallel for num_threads(threads) schedule(static)
    for (int64_t p = 0; p < np; p++) {
      T Cb[9];
      for (int i = 0; i < 9; i++) Cb[i] = Ck[9 * p + i];
      for (int k = 0; k < 3; k++) Cb[4 * k] = damp(Cd[3 * p + k], mu, ident);
      invert<T, 3>(Cb, &Cinv[9 * p]);
    }
    // execute_schur_multiplication (:649-734): dst -= L * M * R^T; each thread owns a range of block rows
#pragma omp parallel num_threads(threads)
    {
#ifdef _OPENMP
      const int tid = omp_get_thread_num(), nt = omp_get_num_threads();
#else
      const int tid = 0, nt = 1;
#endif
      const int64_t r0 = nc * tid / nt, r1 = nc * (tid + 1) / nt;
      for (int64_t p = 0; p < np; p++) {
        const T *Ci = &Cinv[9 * p];
        for (int64_t a = pptr[p]; a < pptr[p + 1]; a++) {
          const int64_t i = ci[a];
          if (i < r0 || i >= r1) continue;
          // LM = E_a * Cinv (9x3)
          T LM[27];
          const T *Ea = &Ek[27 * a];
          for (int cc = 0; cc < 3; cc++)
            for (int rr = 0; rr < 9; rr++)
              LM[rr + 9 * cc] = Ea[rr] * Ci[0 + 3 * cc] + Ea[rr + 9] * Ci[1 + 3 * cc] + Ea[rr + 18] * Ci[2 + 3 * cc];
          for (int64_t bb = a; bb < pptr[p + 1]; bb++) {
            const T *Eb = &Ek[27 * bb];
            T *d = &Sv[81 * sfind(i, ci[bb])];
            for (int cc = 0; cc < 9; cc++)
              for (int rr = 0; rr < 9; rr++)
                d[rr + 9 * cc] -= LM[rr] * Eb[cc] + LM[rr + 9] * Eb[cc + 9] + LM[rr + 18] * Eb[cc + 18];
          }
          // b_S contribution (:901-920)
          const T *bl = &b[dimc + 3 * p];
          T *o = &bS[9 * i];
          for (int rr = 0; rr < 9; rr++) o[rr] -= LM[rr] * bl[0] + LM[rr + 9] * bl[1] + LM[rr + 18] * bl[2];
        }
      }
    }
    for (int64_t i = 0; i < dimc; i++) bS[i] += b[i];
    // block-Jacobi: inverse of diag blocks of S (block_jacobi_schur.hpp:114-151)
    Minv.resize(81 * nc);
#pragma omp parallel for num_threads(threads) schedule(static)
    for (int64_t c = 0; c < nc; c++) invert<T, 9>(&Sv[81 * sfind(c, (int32_t)c)], &Minv[81 * c]);
    tim[2] += secs(t0);
  }

  void schur(double mu, int ident, double *Sd, double *bs) override {
    build_hessian();
    build_schur((T)mu, ident != 0);
    if (Sd) {
      std::fill(Sd, Sd + dimc * dimc, 0.0);
      for (int64_t i = 0; i < nc; i++)
        for (int64_t k = srow[i]; k < srow[i + 1]; k++) {
          const int64_t j = scol[k];
          for (int cc = 0; cc < 9; cc++)
            for (int rr = 0; rr < 9; rr++) Sd[(9 * i + rr) + (9 * j + cc) * dimc] = Sv[81 * k + rr + 9 * cc];
        }
    }
    if (bs) for (int64_t i = 0; i < dimc; i++) bs[i] = bS[i];
  }

  // schur.hpp:347-393: y = S x using the upper blocks and their transposes
  void spmv(const T *x, T *y, std::vector<std::vector<T>> &tl) {
#pragma omp parallel num_threads(threads)
    {
#ifdef _OPENMP
      const int tid = omp_get_thread_num();
#else
      const int tid = 0;
#endif
      T *yt = tl[tid].data();
      std::fill(yt, yt + dimc, T(0));
#pragma omp for schedule(dynamic, 16)
      for (int64_t i = 0; i < nc; i++) {
        for (int64_t k = srow[i]; k < srow[i + 1]; k++) {
          const int64_t j = scol[k];
          const T *A = &Sv[81 * k];
          for (int cc = 0; cc < 9; cc++) {
            const T xv = x[9 * j + cc];
            for (int rr = 0; rr < 9; rr++) yt[9 * i + rr] += A[rr + 9 * cc] * xv;
          }
          if (j != i)
            for (int cc = 0; cc < 9; cc++) {
              T acc = 0;
              for (int rr = 0; rr < 9; rr++) acc += A[rr + 9 * cc] * x[9 * i + rr];
              yt[9 * j + cc] += acc;
            }
        }
      }
    }
    std::fill(y, y + dimc, T(0));
    for (int t = 0; t < threads; t++)
      for (int64_t i = 0; i < dimc; i++) y[i] += tl[t][i];
  }
  void precond(const T *rr, T *z) {
#pragma omp parallel for num_threads(threads) schedule(static)
    for (int64_t c = 0; c < nc; c++)
      for (int row = 0; row < 9; row++) {
        T acc = 0;
        for (int k = 0; k < 9; k++) acc += Minv[81 * c + row + 9 * k] * rr[9 * c + k];
        z[9 * c + row] = acc;
      }
  }
  static T dot(const T *a, const T *bb, int64_t n) {
    T s = 0;
    for (int64_t i = 0; i < n; i++) s += a[i] * bb[i];
    return s;
  }

  // pcg_schur.hpp:79-168
  int64_t pcg(const orc_lm_options *o, T *x) {
    auto t0 = clk::now();
    std::vector<T> rv(bS), z(dimc), p(dimc), Ap(dimc), xb(dimc);
    std::vector<std::vector<T>> tl(threads, std::vector<T>(dimc));
    std::fill(x, x + dimH, T(0));
    precond(rv.data(), z.data());
    p = z;
    T rz = dot(rv.data(), z.data(), dimc);
    T rz0 = std::numeric_limits<T>::infinity();
    const T tol = (T)o->pcg_tolerance, ratio = (T)o->rejection_ratio;
    int64_t k = 0, done = 0;
    for (; k < o->pcg_iterations; ++k) {
      if (rz == T(0)) break;
      spmv(p.data(), Ap.data(), tl);
      const T denom = dot(p.data(), Ap.data(), dimc);
      if (denom == T(0) || std::isnan(denom)) break;
      const T alpha = rz / denom;
      for (int64_t i = 0; i < dimc; i++) xb[i] = x[i];
      for (int64_t i = 0; i < dimc; i++) x[i] = alpha * p[i] + x[i];
      for (int64_t i = 0; i < dimc; i++) rv[i] = -alpha * Ap[i] + rv[i];
      precond(rv.data(), z.data());
      const T rzn = dot(rv.data(), z.data(), dimc);
      done = k + 1;
      if (std::abs(rzn) > ratio * rz0 || std::isnan(rzn)) {
        for (int64_t i = 0; i < dimc; i++) x[i] = xb[i];
        break;
      }
      rz0 = std::min(rz0, std::abs(rzn));
      const T beta = rzn / rz;
      rz = rzn;
      for (int64_t i = 0; i < dimc; i++) p[i] = beta * p[i] + z[i];
      if (std::abs(rzn) < tol) break;
    }
    tim[3] += secs(t0);
    return done;
  }

  // eigen_schur.hpp:52-108 / src/eigen_solver.cpp:10-29 restated as dense LDL^T of S (upper)
  bool direct(T *x) {
    auto t0 = clk::now();
    const int64_t n = dimc;
    std::vector<double> A((size_t)n * n, 0.0);
    for (int64_t i = 0; i < nc; i++)
      for (int64_t k = srow[i]; k < srow[i + 1]; k++) {
        const int64_t j = scol[k];
        for (int cc = 0; cc < 9; cc++)
          for (int rr = 0; rr < 9; rr++) {
            const double v = Sv[81 * k + rr + 9 * cc];
            if (9 * i + rr <= 9 * j + cc) { A[(9 * i + rr) * n + 9 * j + cc] = v; A[(9 * j + cc) * n + 9 * i + rr] = v; }
          }
      }
    // in-place LDL^T (lower), no pivoting
    std::vector<double> D(n);
    for (int64_t j = 0; j < n; j++) {
      double d = A[j * n + j];
      for (int64_t k = 0; k < j; k++) d -= A[j * n + k] * A[j * n + k] * D[k];
      D[j] = d;
      if (d == 0.0 || !std::isfinite(d)) return false;
#pragma omp parallel for num_threads(threads) schedule(static)
      for (int64_t i = j + 1; i < n; i++) {
        double v = A[i * n + j];
        for (int64_t k = 0; k < j; k++) v -= A[i * n + k] * A[j * n + k] * D[k];
        A[i * n + j] = v / d;
      }
    }
    std::vector<double> y(n);
    for (int64_t i = 0; i < n; i++) { double v = bS[i]; for (int64_t k = 0; k < i; k++) v -= A[i * n + k] * y[k]; y[i] = v; }
    for (int64_t i = 0; i < n; i++) y[i] /= D[i];
    for (int64_t i = n - 1; i >= 0; i--) { double v = y[i]; for (int64_t k = i + 1; k < n; k++) v -= A[k * n + i] * y[k]; y[i] = v; }
    std::fill(x, x + dimH, T(0));
    for (int64_t i = 0; i < n; i++) x[i] = (T)y[i];
    tim[3] += secs(t0);
    return true;
  }

  // schur.hpp:279-302: x_l = Hll^-1 (b_l - Hpl^T x_p)
  void backsubst(T *x) {
    auto t0 = clk::now();
#pragma omp parallel for num_threads(threads) schedule(static)
    for (int64_t p = 0; p < np; p++) {
      T rhs[3] = {0, 0, 0};
      for (int64_t f = pptr[p]; f < pptr[p + 1]; f++) {
        const T *Eb = &Ek[27 * f];
        const T *xc = &x[9 * ci[f]];
        for (int cc = 0; cc < 3; cc++) {
          T acc = 0;
          for (int rr = 0; rr < 9; rr++) acc += Eb[rr + 9 * cc] * xc[rr];
          rhs[cc] += acc;
        }
      }
      for (int cc = 0; cc < 3; cc++) rhs[cc] = T(-1) * rhs[cc] + b[dimc + 3 * p + cc];
      const T *Ci = &Cinv[9 * p];
      for (int rr = 0; rr < 3; rr++) x[dimc + 3 * p + rr] = Ci[rr] * rhs[0] + Ci[rr + 3] * rhs[1] + Ci[rr + 6] * rhs[2];
    }
    tim[4] += secs(t0);
  }

  // solver/pcg.hpp:61-232 with BlockJacobiPreconditioner (preconditioner/block_jacobi.hpp:79-186): matrix-free PCG on
  // the FULL system (cameras and points), operator J~^T J~ + mu clamp(diag), preconditioner = inverse of the damped
  // per-vertex diagonal blocks, applied to the NORMALISED residual y = r / |r|.
  int64_t pcg_full(const orc_lm_options *o, T mu, T *x) {
    auto t0 = clk::now();
    const bool ident = o->use_identity != 0;
    // diag of J~^T J~ (scaled Jacobians), clamped (pcg.hpp:93-104)
    std::vector<T> diag(dimH);
    for (int64_t i = 0; i < dimc; i++) diag[i] = Bd[i];
    for (int64_t i = 0; i < 3 * np; i++) diag[dimc + i] = Cd[i];
    for (auto &d : diag) d = std::min(std::max(d, (T)1.0e-6), (T)1.0e32);
    // preconditioner blocks: diag_i + mu clamp(diag_i) on the diagonal, computed in double (ops/hessian.hpp:80-110)
    std::vector<T> Mc(81 * nc), Mp(9 * np);
#pragma omp parallel for num_threads(threads) schedule(static)
    for (int64_t c = 0; c < nc; c++) {
      T blk[81];
      for (int i = 0; i < 81; i++) blk[i] = Bk[81 * c + i];
      for (int k = 0; k < 9; k++) blk[10 * k] = damp(Bd[9 * c + k], mu, ident);
      invert<T, 9>(blk, &Mc[81 * c]);
    }
#pragma omp parallel for num_threads(threads) schedule(static)
    for (int64_t p = 0; p < np; p++) {
      T blk[9];
      for (int i = 0; i < 9; i++) blk[i] = Ck[9 * p + i];
      for (int k = 0; k < 3; k++) blk[4 * k] = damp(Cd[3 * p + k], mu, ident);
      invert<T, 3>(blk, &Mp[9 * p]);
    }
    auto apply = [&](const T *in, T *out) {
#pragma omp parallel for num_threads(threads) schedule(static)
      for (int64_t c = 0; c < nc; c++)
        for (int row = 0; row < 9; row++) {
          T acc = 0;
          for (int k = 0; k < 9; k++) acc += Mc[81 * c + row + 9 * k] * in[9 * c + k];
          out[9 * c + row] = acc;
        }
#pragma omp parallel for num_threads(threads) schedule(static)
      for (int64_t p = 0; p < np; p++)
        for (int row = 0; row < 3; row++) {
          T acc = 0;
          for (int k = 0; k < 3; k++) acc += Mp[9 * p + row + 3 * k] * in[dimc + 3 * p + k];
          out[dimc + 3 * p + row] = acc;
        }
    };
    // v2 = J~^T (J~ p) (ops/product.hpp:49-99, 226-288), then += mu diag p (ops/vector.hpp:24-39)
    std::vector<T> v1(2 * m), v2(dimH);
    std::vector<std::vector<T>> tl(threads, std::vector<T>(dimc));
    auto op = [&](const T *pv) {
#pragma omp parallel for num_threads(threads) schedule(static)
      for (int64_t f = 0; f < m; f++) {
        const T *J = &Jc[18 * f], *Q = &Jp[6 * f];
        const T *pc = pv + 9 * ci[f], *pp = pv + dimc + 3 * pi[f];
        T a0 = 0, a1 = 0;
        for (int k = 0; k < 9; k++) { a0 += J[2 * k] * pc[k]; a1 += J[2 * k + 1] * pc[k]; }
        for (int k = 0; k < 3; k++) { a0 += Q[2 * k] * pp[k]; a1 += Q[2 * k + 1] * pp[k]; }
        v1[2 * f] = a0; v1[2 * f + 1] = a1;
      }
#pragma omp parallel num_threads(threads)
      {
#ifdef _OPENMP
        const int tid = omp_get_thread_num();
#else
        const int tid = 0;
#endif
        T *d = tl[tid].data();
        std::fill(d, d + dimc, T(0));
#pragma omp for schedule(static)
        for (int64_t f = 0; f < m; f++) {
          const T *J = &Jc[18 * f];
          T *dd = d + 9 * ci[f];
          for (int k = 0; k < 9; k++) dd[k] += pw(&J[2 * k], &v1[2 * f], &Pm[4 * f]) * dLv[f]; // product.hpp:270-282
        }
      }
      for (int64_t i = 0; i < dimc; i++) { T a = 0; for (int t = 0; t < threads; t++) a += tl[t][i]; v2[i] = a; }
#pragma omp parallel for num_threads(threads) schedule(static)
      for (int64_t p = 0; p < np; p++) {
        T acc[3] = {0, 0, 0};
        for (int64_t f = pptr[p]; f < pptr[p + 1]; f++) {
          const T *Q = &Jp[6 * f];
          for (int k = 0; k < 3; k++) acc[k] += pw(&Q[2 * k], &v1[2 * f], &Pm[4 * f]) * dLv[f];
        }
        for (int k = 0; k < 3; k++) v2[dimc + 3 * p + k] = acc[k];
      }
      for (int64_t i = 0; i < dimH; i++) v2[i] += ident ? mu * pv[i] : mu * diag[i] * pv[i];
    };
    std::vector<T> rv(b), y(dimH), z(dimH), p(dimH), xb(dimH);
    std::fill(x, x + dimH, T(0));
    T rnorm = std::sqrt(dot(rv.data(), rv.data(), dimH));
    for (int64_t i = 0; i < dimH; i++) y[i] = (T)(1.0 / rnorm) * rv[i];
    apply(y.data(), z.data());
    p = z;
    T rz = dot(rv.data(), z.data(), dimH);
    T rz0 = std::numeric_limits<T>::infinity();
    const T tol = (T)o->pcg_tolerance, ratio = (T)o->rejection_ratio;
    int64_t done = 0;
    for (int64_t k = 0; k < o->pcg_iterations; ++k) {
      if (rz == 0) break;
      op(p.data());
      const T alpha = rz / dot(p.data(), v2.data(), dimH);
      for (int64_t i = 0; i < dimH; i++) xb[i] = x[i];
      for (int64_t i = 0; i < dimH; i++) x[i] = alpha * p[i] + x[i];
      for (int64_t i = 0; i < dimH; i++) rv[i] = -alpha * v2[i] + rv[i];
      rnorm = std::sqrt(dot(rv.data(), rv.data(), dimH));
      for (int64_t i = 0; i < dimH; i++) y[i] = (T)(1.0 / rnorm) * rv[i];
      apply(y.data(), z.data());
      const T rzn = dot(rv.data(), z.data(), dimH);
      done = k + 1;
      if (std::abs(rzn) > ratio * rz0 || std::isnan(rzn)) {
        for (int64_t i = 0; i < dimH; i++) x[i] = xb[i];
        break;
      }
      rz0 = std::min(rz0, std::abs(rzn));
      const T beta = rzn / rz;
      rz = rzn;
      for (int64_t i = 0; i < dimH; i++) p[i] = beta * p[i] + z[i];
      if (std::abs(rzn) < tol) break;
    }
    tim[3] += secs(t0);
    return done;
  }

  int64_t do_solve(const orc_lm_options *o, T mu, T *x, bool &ok) {
    if (o->solver == 2) { ok = true; return pcg_full(o, mu, x); }
    build_schur(mu, o->use_identity != 0);
    int64_t k = 0;
    ok = true;
    if (o->solver == 1) ok = direct(x); else k = pcg(o, x);
    if (ok) backsubst(x);
    return k;
  }
  int64_t solve(const orc_lm_options *o, double mu, double *delta) override {
    set_threads(o->threads);
    build_hessian();
    std::vector<T> x(dimH);
    bool ok;
    int64_t k = do_solve(o, (T)mu, x.data(), ok);
    for (int64_t i = 0; i < dimH; i++) delta[i] = x[i];
    return ok ? k : -1;
  }

  // levenberg_marquardt.hpp:109-242, split into "everything before the loop" and "one loop body"
  T mu_s = 0, nu_s = 2, chi2_s = 0;
  bool run_s = true;
  void lm_begin(const orc_lm_options *o) override {
    set_threads(o->threads);
    for (double &t : tim) t = 0;
    mu_s = (T)o->initial_damping;
    nu_s = 2;
    do_linearize();
    build_hessian();
    chi2_s = compute_error_chi2();
    run_s = true;
  }
  // returns 1 to continue, 0 when the loop would terminate after this iteration
  int lm_step(const orc_lm_options *o, double *out4) override {
    std::vector<T> dx(dimH);
    bool ok;
    const int64_t k = do_solve(o, mu_s, dx.data(), ok);
    auto t0 = clk::now();
    cams_bak = cams; pts_bak = pts;
    // ops/update.hpp:23-30
    for (int64_t i = 0; i < dimc; i++) cams[i] += dx[i] * scales[i];
    for (int64_t i = 0; i < 3 * np; i++) pts[i] += dx[dimc + i] * scales[dimc + i];
    T nchi2 = compute_error_chi2();
    if (!ok) nchi2 = std::numeric_limits<T>::max();
    // compute_rho :19-47
    T num = chi2_s - nchi2, denom = 1.0;
    if (ok) {
      T sacc = 0;
      for (int64_t i = 0; i < dimH; i++) sacc += dx[i] * (mu_s * dx[i] + b[i]);
      denom = sacc + (T)1.0e-3;
    }
    const T rho = num / denom;
    tim[5] += secs(t0);
    if (ok && std::isfinite(nchi2) && rho > 0) {
      double alpha = 1.0 - std::pow(2.0 * rho - 1.0, 3);
      alpha = std::max(std::min(alpha, 2.0 / 3.0), 1.0 / 3.0);
      mu_s *= (T)alpha;
      nu_s = 2;
      do_linearize();
      build_hessian();
    } else {
      cams = cams_bak; pts = pts_bak;
      compute_error_chi2();
      mu_s *= nu_s;
      nu_s *= 2;
      nchi2 = chi2_s;
    }
    out4[0] = chi2_s; out4[1] = nchi2; out4[2] = mu_s; out4[3] = (double)k;
    chi2_s = nchi2;
    if (!std::isfinite(mu_s)) run_s = false;
    if (rho == 0) return 0;
    return run_s ? 1 : 0;
  }
  int64_t lm(const orc_lm_options *o, double *traj) override {
    lm_begin(o);
    int64_t it = 0;
    while (it < o->iterations) {
      const int go = lm_step(o, traj + 4 * it);
      it++;
      if (!go) break;
    }
    return it;
  }
};

} // namespace

struct orc_problem { Base *impl; };

extern "C" {
orc_problem *orc_create_f64(int64_t nc, int64_t np, int64_t m, const int32_t *c, const int32_t *p, const double *o,
                            const double *cm, const double *pt) {
  return new orc_problem{new Impl<double>(nc, np, m, c, p, o, cm, pt)};
}
orc_problem *orc_create_f32(int64_t nc, int64_t np, int64_t m, const int32_t *c, const int32_t *p, const double *o,
                            const double *cm, const double *pt) {
  return new orc_problem{new Impl<float>(nc, np, m, c, p, o, cm, pt)};
}
void orc_destroy(orc_problem *h) { if (h) { delete h->impl; delete h; } }
void orc_set_threads(orc_problem *h, int t) { h->impl->set_threads(t); }
void orc_get_params(orc_problem *h, double *c, double *p) { h->impl->get_params(c, p); }
void orc_set_params(orc_problem *h, const double *c, const double *p) { h->impl->set_params(c, p); }
void orc_set_robust(orc_problem *h, int loss_kind, double delta, const double *P) { h->impl->set_robust(loss_kind, delta, P); }
double orc_residuals(orc_problem *h, double *r) { return h->impl->residuals(r); }
void orc_jacobians(orc_problem *h, double *a, double *b) { h->impl->jacobians(a, b); }
double orc_linearize(orc_problem *h, double *s, double *b) { return h->impl->linearize(s, b); }
void orc_hessian_structure(orc_problem *h, int64_t *a, int64_t *b, int64_t *c) { h->impl->hessian_structure(a, b, c); }
int64_t orc_hessian_num_values(orc_problem *h) { return h->impl->hessian_num_values(); }
void orc_hessian_values(orc_problem *h, double *v) { h->impl->hessian_values(v); }
void orc_schur(orc_problem *h, double mu, int id, double *S, double *bS) { h->impl->schur(mu, id, S, bS); }
int64_t orc_schur_nnz_blocks(orc_problem *h) { return h->impl->schur_nnz_blocks(); }
int64_t orc_solve(orc_problem *h, const orc_lm_options *o, double mu, double *d) { return h->impl->solve(o, mu, d); }
int64_t orc_lm(orc_problem *h, const orc_lm_options *o, double *t) { return h->impl->lm(o, t); }
void orc_lm_begin(orc_problem *h, const orc_lm_options *o) { h->impl->lm_begin(o); }
int orc_lm_step(orc_problem *h, const orc_lm_options *o, double *out4) { return h->impl->lm_step(o, out4); }
void orc_last_timings(orc_problem *h, double *t) { for (int i = 0; i < 6; i++) t[i] = h->impl->tim[i]; }
}

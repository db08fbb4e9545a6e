// graph_generic.cu — the generic factor-graph path behind include/graphite_b200_graph.h.
//
// Vertex sets of any dimension, factor sets of any arity / residual size, fixed vertices, activity levels, precision
// matrices, robust loss; factors are evaluated by the caller's kernels (gb_graph_factor_fn), everything after that —
// chi2 / loss weights, Jacobi scales, gradient, Hessian blocks, J v / J^T P v, block-Jacobi PCG on the full system, step,
// rho, the LM loop — runs here.  Reference: Graph (graph.hpp:92-318), FactorDescriptor (factor.hpp), ops/*.hpp,
// PCGSolver (solver/pcg.hpp:61-232), BlockJacobiPreconditioner (preconditioner/block_jacobi.hpp:79-186),
// levenberg_marquardt (optimizer/levenberg_marquardt.hpp:109-242).
//
// Design: every sum over factors is a GATHER over a per-vertex incidence list built once per initialisation (the
// reference scatters with atomicAdd from one thread per output scalar); sums therefore have a fixed order and runs are
// bit-reproducible.  A PCG solve is ONE cooperative kernel (all iterations, scalars on the device, grid barriers between
// the phases) instead of ~15 launches, 3 blocking reductions and 3 stream synchronisations per iteration
// (solver/pcg.hpp:141-222).  Jacobians keep the reference's per-factor E x d column-major blocks: a vertex-side gather
// then reads one contiguous block per incident factor.
#include "context.hpp"
#include "../../include/graphite_b200_graph.h"

#include <algorithm>
#include <cmath>
#include <cooperative_groups.h>
#include <cstring>
#include <cuda_bf16.h>
#include <limits>
#include <map>
#include <vector>

namespace cg = cooperative_groups;

namespace gg {

constexpr int MAX_SETS = 16;

template <typename T, typename S> __device__ __forceinline__ T ld(const S &v) {
  if constexpr (std::is_same<S, __nv_bfloat16>::value) return (T)__bfloat162float(v);
  else return (T)v;
}
template <typename S, typename T> __device__ __forceinline__ S st(const T &v) {
  if constexpr (std::is_same<S, __nv_bfloat16>::value) {
    if constexpr (std::is_same<T, double>::value) return __double2bfloat16(v);
    else return __float2bfloat16(v);
  } else return (S)v;
}
// InvP<T,S> (types.hpp:19-20): the preconditioner blocks are kept in S, or in T when S is a 16-bit type
template <typename T, typename S> struct InvPT { using type = S; };
template <typename T> struct InvPT<T, __nv_bfloat16> { using type = T; };

// ---------------------------------------------------------------- device tables
template <typename T, typename S> struct DFSet {
  int E, arity, loss, pad;
  int d[GB_MAX_ARITY], vset[GB_MAX_ARITY];
  long long count, nactive, roff;
  double delta;
  const int *vidx;       // [count][arity]
  const int *active_idx; // [nactive]
  T *r;                  // [count][E]   residuals at the linearisation point
  T *chi2;               // [count]
  S *dL;                 // [count]
  S *P;                  // [count][E*E]
  S *J[GB_MAX_ARITY];    // [count][E*d]
};
template <typename T> struct DVSet {
  int dim, npar;
  long long count;
  T *params;
  const unsigned char *active; // 1 = active (has a Hessian column)
  const int *hoff;             // hessian column offset, -1 inactive
  const long long *inc_ptr;    // [count+1]
  const unsigned long long *inc; // fset:8 | slot:8 | pad:16 | factor:32
  void *blockdiag, *pinv;      // [count][d*d] of InvP
  T *sdiag;                    // [count][d] undamped diagonal (block_jacobi.hpp:98-112)
};

__device__ __forceinline__ void unpack_inc(unsigned long long e, int &fs, int &slot, long long &f) {
  fs = (int)(e >> 56);
  slot = (int)((e >> 48) & 0xff);
  f = (long long)(e & 0xffffffffull);
}

// deterministic block sum (fixed tree), result valid in thread 0
template <typename T> __device__ T block_sum(T v, T *sh) {
  const int t = threadIdx.x;
  sh[t] = v;
  __syncthreads();
  for (int s = blockDim.x >> 1; s > 0; s >>= 1) {
    if (t < s) sh[t] += sh[t + s];
    __syncthreads();
  }
  T out = sh[0];
  __syncthreads();
  return out;
}

// ---------------------------------------------------------------- linearisation kernels
// chi2_f = loss(r^T P r), dL = loss'(r^T P r) (ops/chi2.hpp:9-44, loss.hpp:20-50); per-block partial sums of chi2_f
template <typename T, typename S>
__global__ void k_g_chi2(DFSet<T, S> F, const T *r, T *chi2_out, S *dL_out, T *partial) {
  extern __shared__ unsigned char smem_raw[];
  T *sh = (T *)smem_raw;
  T acc = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < F.nactive; i += (long long)gridDim.x * blockDim.x) {
    const long long f = F.active_idx[i];
    const int E = F.E;
    const T *rf = r + f * E;
    const S *P = F.P + f * E * E;
    T c = 0;
    for (int a = 0; a < E; a++) {
      T pr = 0;
      for (int b = 0; b < E; b++) pr += ld<T>(P[a * E + b]) * rf[b];
      c += pr * rf[a];
    }
    T val = c, der = T(1);
    if (F.loss == GB_LOSS_HUBER) {
      const T dl = (T)F.delta;
      if (c > dl * dl) {
        const T sq = sqrt(c);
        val = T(2) * sq * dl - dl * dl;
        der = dl / sq;
      }
    }
    if (chi2_out) chi2_out[f] = val;
    if (dL_out) dL_out[f] = st<S>(der);
    acc += val;
  }
  const T s = block_sum(acc, sh);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}
// out[slot] (+)= sum of partial[0..n) in order (one thread: n <= a few hundred)
template <typename T> __global__ void k_g_sum(const T *partial, int n, T *out, int accumulate) {
  T s = 0;
  for (int i = 0; i < n; i++) s += partial[i];
  *out = accumulate ? *out + s : s;
}

// cast the caller's T Jacobians into the S store; slots of inactive vertices and inactive factors are zero
// (ops/linearize.hpp:127-129 fills with 0, :24-27 skips inactive vertices)
template <typename T, typename S>
__global__ void k_g_store_jac(DFSet<T, S> F, int slot, const T *src, const unsigned char *factive, const unsigned char *vactive) {
  const int ED = F.E * F.d[slot];
  const long long n = F.count * ED;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long f = i / ED;
    const bool on = factive[f] && vactive[F.vidx[f * F.arity + slot]];
    F.J[slot][i] = on ? st<S>(src[i]) : st<S>(T(0));
  }
}
// J[:, col] <- (S)((T)J[:, col] * scale) (ops/linearize.hpp:140-180)
template <typename T, typename S>
__global__ void k_g_scale_jac(DFSet<T, S> F, int slot, const T *scales, const int *hoff) {
  const int D = F.d[slot], E = F.E;
  const long long n = F.nactive * D;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long f = F.active_idx[i / D];
    const int col = (int)(i % D);
    const int ho = hoff[F.vidx[f * F.arity + slot]];
    if (ho < 0) continue;
    const T s = scales[ho + col];
    S *J = F.J[slot] + f * E * D + col * E;
    for (int a = 0; a < E; a++) J[a] = st<S>(ld<T>(J[a]) * s);
  }
}

// one thread per Hessian scalar column; gathers over the incident factors of the column's vertex.
// MODE 0: diag_c = sum dL (J^T P J)_cc (ops/hessian.hpp:418-474)   MODE 1: b_c = -sum J_c^T (dL P r) (ops/linearize.hpp:238-303)
// MODE 2: y_c = sum dL J_c^T P v (ops/product.hpp:228-290)
template <typename T, typename S, int MODE>
__device__ __forceinline__ T gather_column(const DFSet<T, S> *fs, const DVSet<T> *vs, unsigned long long cm, const T *vec) {
  const int vsi = (int)(cm >> 40);
  const long long v = (long long)((cm >> 8) & 0xffffffffull);
  const int col = (int)(cm & 0xff);
  const DVSet<T> &V = vs[vsi];
  T acc = 0;
  for (long long e = V.inc_ptr[v]; e < V.inc_ptr[v + 1]; e++) {
    int fsi, slot;
    long long f;
    unpack_inc(V.inc[e], fsi, slot, f);
    const DFSet<T, S> &F = fs[fsi];
    const int E = F.E;
    const S *jcol = F.J[slot] + (f * F.d[slot] + col) * E;
    const S *P = F.P + f * E * E;
    const T dL = ld<T>(F.dL[f]);
    T value = 0;
    if (MODE == 0) {
      for (int a = 0; a < E; a++) {
        T pj = 0;
        for (int b = 0; b < E; b++) pj += ld<T>(P[a * E + b]) * ld<T>(jcol[b]);
        value += ld<T>(jcol[a]) * pj;
      }
      value *= dL;
    } else if (MODE == 1) {
      const T *r = F.r + f * E;
      for (int a = 0; a < E; a++) {
        T x2 = 0;
        for (int b = 0; b < E; b++) x2 += dL * ld<T>(P[a * E + b]) * r[b];
        value -= ld<T>(jcol[a]) * x2;
      }
    } else {
      const T *x = vec + F.roff + f * E;
      for (int a = 0; a < E; a++) {
        T x2 = 0;
        for (int b = 0; b < E; b++) x2 += ld<T>(P[a * E + b]) * x[b];
        value += ld<T>(jcol[a]) * x2;
      }
      value *= dL;
    }
    acc += value;
  }
  return acc;
}
template <typename T, typename S, int MODE>
__global__ void k_g_columns(const DFSet<T, S> *fs, const DVSet<T> *vs, const unsigned long long *colmap, long long dimH,
                            const T *vec, T *out, int to_scale) {
  for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < dimH; c += (long long)gridDim.x * blockDim.x) {
    T v = gather_column<T, S, MODE>(fs, vs, colmap[c], vec);
    if (to_scale) v = (T)(1.0 / (2.220446049250313e-16 + sqrt((double)v))); // graph.hpp:262-270
    out[c] = v;
  }
}

// y[row] = sum over the factor's active vertices of J[row, :] x[vertex] (ops/product.hpp:51-99)
template <typename T, typename S>
__device__ __forceinline__ void jv_rows(const DFSet<T, S> &F, const DVSet<T> *vs, const T *x, T *y, long long start, long long stride) {
  const int E = F.E;
  for (long long i = start; i < F.nactive * E; i += stride) {
    const long long f = F.active_idx[i / E];
    const int row = (int)(i % E);
    T value = 0;
    for (int s = 0; s < F.arity; s++) {
      const int ho = vs[F.vset[s]].hoff[F.vidx[f * F.arity + s]];
      if (ho < 0) continue;
      const int D = F.d[s];
      const S *jrow = F.J[s] + f * E * D + row;
      T part = 0;
      for (int k = 0; k < D; k++) part += ld<T>(jrow[k * E]) * x[ho + k];
      value += part;
    }
    y[F.roff + f * E + row] = value;
  }
}
template <typename T, typename S>
__global__ void k_g_jv(const DFSet<T, S> *fs, int nf, const DVSet<T> *vs, const T *x, T *y) {
  for (int i = 0; i < nf; i++)
    jv_rows<T, S>(fs[i], vs, x, y, blockIdx.x * (long long)blockDim.x + threadIdx.x, (long long)gridDim.x * blockDim.x);
}

// block diagonal of J~^T dL P J~ per vertex, column-major d x d (ops/hessian.hpp:166-260), and its scalar diagonal
template <typename T, typename S, typename IP>
__global__ void k_g_blockdiag(const DFSet<T, S> *fs, const DVSet<T> *vs, int vsi) {
  const DVSet<T> V = vs[vsi];
  const int D = V.dim, BS = D * D;
  IP *blocks = (IP *)V.blockdiag;
  const long long n = V.count * BS;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long v = i / BS;
    const int off = (int)(i % BS), row = off % D, col = off / D;
    T acc = 0;
    if (V.active[v]) {
      for (long long e = V.inc_ptr[v]; e < V.inc_ptr[v + 1]; e++) {
        int fsi, slot;
        long long f;
        unpack_inc(V.inc[e], fsi, slot, f);
        const DFSet<T, S> &F = fs[fsi];
        const int E = F.E;
        const S *Jt = F.J[slot] + (f * D + row) * E, *J = F.J[slot] + (f * D + col) * E;
        const S *P = F.P + f * E * E;
        T value = 0;
        for (int a = 0; a < E; a++) {
          T pj = 0;
          for (int b = 0; b < E; b++) pj += ld<T>(P[a * E + b]) * ld<T>(J[b]);
          value += ld<T>(Jt[a]) * pj;
        }
        acc += value * ld<T>(F.dL[f]);
      }
    }
    blocks[i] = (IP)acc;
    if (row == col) V.sdiag[v * D + col] = (T)(IP)acc;
  }
}

// damp the diagonal from the backed-up values (ops/hessian.hpp:80-109) and invert the block (the reference: cuBLAS
// matinvBatched, block_jacobi.hpp:141-165): Gauss-Jordan with partial pivoting in IP, one thread per vertex
template <typename T, typename IP>
__global__ void k_g_invert(const DVSet<T> *vs, int vsi, double mu, int use_identity) {
  const DVSet<T> V = vs[vsi];
  const int D = V.dim, BS = D * D;
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < V.count; v += (long long)gridDim.x * blockDim.x) {
    if (!V.active[v]) continue;
    IP *blk = (IP *)V.blockdiag + v * BS;
    IP *inv = (IP *)V.pinv + v * BS;
    IP a[GB_MAX_DIM * GB_MAX_DIM], b[GB_MAX_DIM * GB_MAX_DIM];
    for (int i = 0; i < BS; i++) { a[i] = blk[i]; b[i] = IP(0); }
    for (int i = 0; i < D; i++) {
      const double dg = (double)V.sdiag[v * D + i];
      const double nd = use_identity ? dg + (double)(IP)mu : dg + (double)(IP)mu * fmin(fmax(dg, 1.0e-6), 1.0e32);
      a[i * D + i] = (IP)nd;
      blk[i * D + i] = (IP)nd; // the reference damps the stored block in place (augment_hessian_diagonal_kernel)
      b[i * D + i] = IP(1);
    }
    for (int k = 0; k < D; k++) {
      int piv = k;
      IP best = fabs(a[k + k * D]);
      for (int r = k + 1; r < D; r++) {
        const IP c = fabs(a[r + k * D]);
        if (c > best) { best = c; piv = r; }
      }
      if (piv != k)
        for (int c = 0; c < D; c++) {
          IP t = a[k + c * D]; a[k + c * D] = a[piv + c * D]; a[piv + c * D] = t;
          t = b[k + c * D]; b[k + c * D] = b[piv + c * D]; b[piv + c * D] = t;
        }
      const IP ip = IP(1) / a[k + k * D];
      for (int c = 0; c < D; c++) { a[k + c * D] *= ip; b[k + c * D] *= ip; }
      for (int r = 0; r < D; r++) {
        if (r == k) continue;
        const IP m = a[r + k * D];
        if (m == IP(0)) continue;
        for (int c = 0; c < D; c++) { a[r + c * D] -= m * a[k + c * D]; b[r + c * D] -= m * b[k + c * D]; }
      }
    }
    for (int i = 0; i < BS; i++) inv[i] = b[i];
  }
}

// ---------------------------------------------------------------- PCG (one cooperative kernel per solve)
struct PcgState {
  long long iterations;
  double rz_final;
  int stop_reason, pad;
};
template <typename T> struct PcgArgs {
  long long dimH, rows;
  int nf, nv, max_iter, use_identity;
  T mu, tol, ratio;
  const T *b, *cdiag; // clamped scalar diagonal of the scaled system (pcg.hpp:93-104)
  T *x, *xb, *r, *z, *p, *v1, *v2;
  T *partial; // [3][gridDim]
  const unsigned long long *colmap;
  PcgState *state;
};

// sum of the per-block partials in block order: every block computes the same value
template <typename T> __device__ T grid_total(const T *partial, int nblocks, T *sh) {
  if (threadIdx.x == 0) {
    T s = 0;
    for (int i = 0; i < nblocks; i++) s += partial[i];
    sh[0] = s;
  }
  __syncthreads();
  const T out = sh[0];
  __syncthreads();
  return out;
}

// z_c = sum_i Pinv[row, i] * (r_i * inv_norm) for the vertex of column c (ops/hessian.hpp:128-150; pcg.hpp:108-120)
template <typename T, typename IP>
__device__ __forceinline__ T precond_col(const DVSet<T> *vs, unsigned long long cm, const T *r, T inv_norm) {
  const DVSet<T> &V = vs[(int)(cm >> 40)];
  const long long v = (long long)((cm >> 8) & 0xffffffffull);
  const int row = (int)(cm & 0xff), D = V.dim;
  const IP *blk = (const IP *)V.pinv + v * D * D;
  const int ho = V.hoff[v];
  T value = 0;
  for (int i = 0; i < D; i++) value += (T)blk[row + i * D] * (inv_norm * r[ho + i]);
  return value;
}

template <typename T, typename S, typename IP>
__global__ void __launch_bounds__(256) k_g_pcg(const DFSet<T, S> *fs, const DVSet<T> *vs, PcgArgs<T> A) {
  cg::grid_group grid = cg::this_grid();
  __shared__ T sh[256];
  const int nb = gridDim.x;
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x, nth = (long long)nb * blockDim.x;
  T *part0 = A.partial, *part1 = A.partial + nb, *part2 = A.partial + 2 * nb;

  // x = 0, r = b, ||r||
  T acc = 0;
  for (long long c = tid; c < A.dimH; c += nth) {
    const T bc = A.b[c];
    A.x[c] = 0;
    A.r[c] = bc;
    acc += bc * bc;
  }
  T s = block_sum(acc, sh);
  if (threadIdx.x == 0) part0[blockIdx.x] = s;
  grid.sync();
  T rnorm = sqrt(grid_total(part0, nb, sh));
  T scale = (T)(1.0 / rnorm);
  // z = M^-1 (r / ||r||), p = z, rz = r.z
  acc = 0;
  for (long long c = tid; c < A.dimH; c += nth) {
    const T zc = precond_col<T, IP>(vs, A.colmap[c], A.r, scale);
    A.z[c] = zc;
    A.p[c] = zc;
    acc += A.r[c] * zc;
  }
  s = block_sum(acc, sh);
  if (threadIdx.x == 0) part1[blockIdx.x] = s;
  grid.sync();
  T rz = grid_total(part1, nb, sh);
  T rz0 = std::numeric_limits<T>::infinity();
  int reason = 0;
  long long k = 0;
  for (; k < A.max_iter; k++) {
    if (rz == T(0)) { reason = 3; break; }
    // v1 = J p
    for (int i = 0; i < A.nf; i++) jv_rows<T, S>(fs[i], vs, A.p, A.v1, tid, nth);
    grid.sync();
    // v2 = J^T dL P v1 + mu * diag * p ; p.v2
    acc = 0;
    for (long long c = tid; c < A.dimH; c += nth) {
      T v2 = gather_column<T, S, 2>(fs, vs, A.colmap[c], A.v1);
      const T pc = A.p[c];
      v2 += A.use_identity ? A.mu * pc : A.mu * A.cdiag[c] * pc;
      A.v2[c] = v2;
      acc += pc * v2;
    }
    s = block_sum(acc, sh);
    if (threadIdx.x == 0) part0[blockIdx.x] = s;
    grid.sync();
    const T alpha = rz / grid_total(part0, nb, sh);
    // x += alpha p (backup first), r -= alpha v2, ||r||
    acc = 0;
    for (long long c = tid; c < A.dimH; c += nth) {
      const T xc = A.x[c];
      A.xb[c] = xc;
      A.x[c] = alpha * A.p[c] + xc;
      const T rc = -alpha * A.v2[c] + A.r[c];
      A.r[c] = rc;
      acc += rc * rc;
    }
    s = block_sum(acc, sh);
    if (threadIdx.x == 0) part1[blockIdx.x] = s;
    grid.sync();
    rnorm = sqrt(grid_total(part1, nb, sh));
    scale = (T)(1.0 / rnorm);
    acc = 0;
    for (long long c = tid; c < A.dimH; c += nth) {
      const T zc = precond_col<T, IP>(vs, A.colmap[c], A.r, scale);
      A.z[c] = zc;
      acc += A.r[c] * zc;
    }
    s = block_sum(acc, sh);
    if (threadIdx.x == 0) part2[blockIdx.x] = s;
    grid.sync();
    const T rz_new = grid_total(part2, nb, sh);
    if (fabs(rz_new) > A.ratio * rz0 || isnan(rz_new)) {
      for (long long c = tid; c < A.dimH; c += nth) A.x[c] = A.xb[c];
      rz = rz_new;
      reason = 2;
      k++;
      break;
    }
    rz0 = fmin(rz0, fabs(rz_new));
    const T beta = rz_new / rz;
    rz = rz_new;
    for (long long c = tid; c < A.dimH; c += nth) A.p[c] = beta * A.p[c] + A.z[c];
    if (fabs(rz_new) < A.tol) { reason = 1; k++; break; }
    grid.sync();
  }
  if (tid == 0) {
    A.state->iterations = k;
    A.state->rz_final = (double)rz;
    A.state->stop_reason = reason;
  }
}

// ---------------------------------------------------------------- step, rho
// delta = x~ * scale for active vertices (zero otherwise); default update params[0..d) += delta (ops/update.hpp:9-31)
template <typename T>
__global__ void k_g_step(DVSet<T> V, const T *x, const T *scales, T *delta, int apply_default) {
  const int D = V.dim;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < V.count * D; i += (long long)gridDim.x * blockDim.x) {
    const long long v = i / D;
    const int k = (int)(i % D);
    const int ho = V.hoff[v];
    const T dl = ho >= 0 ? x[ho + k] * scales[ho + k] : T(0);
    delta[i] = dl;
    if (apply_default && ho >= 0) V.params[v * V.npar + k] += dl;
  }
}
// sum x (mu x + b) over the Hessian dimension (levenberg_marquardt.hpp:34-43), per-block partials
template <typename T> __global__ void k_g_rho(long long n, const T *x, const T *b, T mu, T *partial) {
  extern __shared__ unsigned char smem_raw[];
  T *sh = (T *)smem_raw;
  T acc = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const T xi = x[i];
    acc += xi * (mu * xi + b[i]);
  }
  const T s = block_sum(acc, sh);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}
template <typename T> __global__ void k_g_clamp(long long n, const T *in, T *out, T lo, T hi) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = fmin(fmax(in[i], lo), hi);
}
template <typename T, typename S> __global__ void k_g_to_T(long long n, const S *in, T *out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) out[i] = ld<T>(in[i]);
}
template <typename T, typename S> __global__ void k_g_from_T(long long n, const T *in, S *out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) out[i] = st<S>(in[i]);
}

// Hessian values: one thread per scalar of a block; contributions (factor, slot_i, slot_j) of the block in a fixed order,
// each rounded to S before it is added, as the reference's atomicAdd of the S-rounded product does (ops/hessian.hpp:58-76)
struct HContrib {
  int fset, si, sj, pad; // block = J[si]^T P J[sj] (si is the row vertex)
  long long f;
};
template <typename T, typename S>
__global__ void k_g_hessian(const DFSet<T, S> *fs, const long long *blk_ptr, const HContrib *contrib, const long long *blk_off,
                            const int *blk_rows, const int *blk_cols, long long nblocks, const long long *scalar_block, long long nvalues, S *H) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvalues; i += (long long)gridDim.x * blockDim.x) {
    const long long bk = scalar_block[i];
    const int off = (int)(i - blk_off[bk]);
    const int dr = blk_rows[bk], row = off % dr, col = off / dr;
    T acc = 0;
    for (long long e = blk_ptr[bk]; e < blk_ptr[bk + 1]; e++) {
      const HContrib c = contrib[e];
      const DFSet<T, S> &F = fs[c.fset];
      const int E = F.E;
      const S *Jt = F.J[c.si] + (c.f * F.d[c.si] + row) * E;
      const S *J = F.J[c.sj] + (c.f * F.d[c.sj] + col) * E;
      const S *P = F.P + c.f * E * E;
      T value = 0;
      for (int a = 0; a < E; a++) {
        T pj = 0;
        for (int b = 0; b < E; b++) pj += ld<T>(P[a * E + b]) * ld<T>(J[b]);
        value += ld<T>(Jt[a]) * pj;
      }
      value *= ld<T>(F.dL[c.f]);
      acc = ld<T>(st<S>(acc + ld<T>(st<S>(value))));
    }
    H[i] = st<S>(acc);
  }
}

// ---------------------------------------------------------------- host side
struct GraphBase {
  gb_context *ctx = nullptr;
  virtual ~GraphBase() {}
  virtual int add_vertex_set(const gb_vertex_set_desc *) = 0;
  virtual int add_factor_set(const gb_factor_set_desc *, gb_graph_factor_fn, void *) = 0;
  virtual int set_update(int, gb_graph_update_fn, void *) = 0;
  virtual int set_fixed(int, const uint8_t *) = 0;
  virtual int set_active(int, const uint8_t *) = 0;
  virtual int set_vertices(int, const void *) = 0;
  virtual int get_vertices(int, void *) = 0;
  virtual int vertices_device(int, void **) = 0;
  virtual int set_precision(int, const void *) = 0;
  virtual int set_loss(int, int, double) = 0;
  virtual int set_scaling(int) = 0;
  virtual int initialize(int, int64_t *) = 0;
  virtual int vertex_columns(int, int64_t *) = 0;
  virtual int hessian_structure(int64_t *, int64_t *, int64_t *) = 0;
  virtual int linearize(double *) = 0;
  virtual int cost(double *) = 0;
  virtual int get(int, int, void *) = 0;
  virtual int hessian_values(void *) = 0;
  virtual int jv(const void *, void *) = 0;
  virtual int jtpv(const void *, void *) = 0;
  virtual int set_damping(double, int) = 0;
  virtual int solve(const gb_pcg_options *, void *, gb_solve_info *) = 0;
  virtual int lm(const gb_lm_options *, gb_lm_result *, double *) = 0;
  virtual int bind_linearization(int, const void *const *, const void *, const void *) = 0;
  virtual int bind_gradient(const void *) = 0;
  virtual int update_values() = 0;
  virtual int solve_device(const gb_pcg_options *, void *, gb_solve_info *) = 0;
};

template <typename T, typename S> struct Graph : GraphBase {
  using IP = typename InvPT<T, S>::type;
  struct VSetH {
    gb_vertex_set_desc d{};
    std::vector<int64_t> gids;
    std::vector<uint8_t> fixed, active; // active: 1 = has a Hessian column at the current level
    std::vector<int64_t> hoff;
    std::vector<int64_t> block;
    gb_graph_update_fn upd = nullptr;
    void *upd_user = nullptr;
    T *params = nullptr, *backup = nullptr, *delta = nullptr, *sdiag = nullptr;
    unsigned char *active_d = nullptr;
    int *hoff_d = nullptr;
    long long *inc_ptr_d = nullptr;
    unsigned long long *inc_d = nullptr;
    IP *blockdiag = nullptr, *pinv = nullptr;
    bool have_values = false;
  };
  struct FSetH {
    gb_factor_set_desc d{};
    std::vector<int32_t> vidx;
    std::vector<uint8_t> level;        // set_active value
    std::vector<uint8_t> active;       // 1 = active at the current level
    std::vector<int32_t> active_idx;
    gb_graph_factor_fn fn = nullptr;
    void *user = nullptr;
    int *vidx_d = nullptr, *active_idx_d = nullptr;
    unsigned char *active_d = nullptr;
    T *r = nullptr, *r_trial = nullptr, *chi2 = nullptr;
    S *dL = nullptr, *P = nullptr;
    S *J[GB_MAX_ARITY] = {nullptr, nullptr, nullptr, nullptr};
    T *Jt[GB_MAX_ARITY] = {nullptr, nullptr, nullptr, nullptr}; // the caller's T output when S != T
    long long roff = 0;
    // a caller-owned linearisation bound in place of the library's buffers (gb_graph_bind_linearization); null = own
    const S *bound_J[GB_MAX_ARITY] = {nullptr, nullptr, nullptr, nullptr};
    const S *bound_dL = nullptr, *bound_P = nullptr;
  };
  std::vector<VSetH> V;
  std::vector<FSetH> F;
  std::vector<void *> struct_allocs; // freed at re-initialisation
  bool initialized = false, linearized = false, damped = false, solved = false, scale_on = true;
  int level = 0;
  long long dimH = 0, nblockcols = 0, rows = 0, nactive_total = 0, bytes = 0;
  // hessian structure (host) + device
  std::vector<int64_t> h_colptr, h_rowidx, h_offsets;
  long long h_nvalues = 0;
  long long *blk_ptr_d = nullptr, *blk_off_d = nullptr, *scalar_block_d = nullptr;
  int *blk_rows_d = nullptr, *blk_cols_d = nullptr;
  HContrib *contrib_d = nullptr;
  // device tables and vectors
  DFSet<T, S> *fs_d = nullptr;
  DVSet<T> *vs_d = nullptr;
  unsigned long long *colmap_d = nullptr;
  T *b = nullptr, *scales = nullptr, *cdiag = nullptr, *x = nullptr, *xb = nullptr, *r = nullptr, *z = nullptr, *p = nullptr,
    *v1 = nullptr, *v2 = nullptr, *partial = nullptr, *scal = nullptr, *tmpH = nullptr;
  PcgState *state_d = nullptr;
  const T *b_bound = nullptr; // caller-owned gradient (gb_graph_bind_gradient)
  bool bound = false;         // the linearisation is the caller's: gb_graph_linearize / gb_graph_lm are not available
  double *h_pin = nullptr; // pinned: [0] chi2 [1] rho denominator; PcgState after it
  double mu = 0;
  int use_identity = 0;
  int nsm = 148, coop_blocks_per_sm = 1;
  static constexpr int NPART = 512;

  ~Graph() override {
    free_structure();
    for (auto &v : V) { cudaFree(v.params); cudaFree(v.backup); }
    for (auto &f : F) cudaFree(f.P);
    if (h_pin) cudaFreeHost(h_pin);
  }
  void free_structure() {
    for (void *a : struct_allocs) cudaFree(a);
    struct_allocs.clear();
    initialized = linearized = damped = solved = false;
  }
  template <typename X> int dalloc(X **out, size_t n) {
    *out = nullptr;
    if (n == 0) n = 1;
    GB_CUDA(ctx, cudaMalloc((void **)out, n * sizeof(X)));
    struct_allocs.push_back((void *)*out);
    bytes += (long long)(n * sizeof(X));
    return GB_OK;
  }
  template <typename X> int upload(X **out, const std::vector<X> &h) {
    GB_TRY(dalloc(out, h.size()));
    if (!h.empty()) GB_CUDA(ctx, cudaMemcpyAsync(*out, h.data(), h.size() * sizeof(X), cudaMemcpyHostToDevice, ctx->stream));
    return GB_OK;
  }
  int require(bool c, const char *what) { return c ? GB_OK : ctx->fail(GB_ERR_INVALID, "%s", what); }
  static int grid_for(long long n) { return (int)std::max<long long>(1, std::min<long long>((n + 255) / 256, 148 * 8)); }
  int launched() {
    GB_LAUNCH(ctx);
    GB_CUDA(ctx, cudaGetLastError());
    return GB_OK;
  }

  int add_vertex_set(const gb_vertex_set_desc *d) override {
    GB_TRY(require(d && d->dimension >= 1 && d->dimension <= GB_MAX_DIM && d->count >= 0 && d->count < (1ll << 31), "bad vertex set"));
    GB_TRY(require((int)V.size() < MAX_SETS, "too many vertex sets"));
    GB_TRY(require(d->global_ids || d->count == 0, "global_ids missing"));
    VSetH v;
    v.d = *d;
    if (v.d.parameters == 0) v.d.parameters = v.d.dimension;
    GB_TRY(require(v.d.parameters >= v.d.dimension, "parameters < dimension"));
    v.gids.assign(d->global_ids, d->global_ids + d->count);
    v.fixed.assign((size_t)d->count, 0);
    if (d->fixed) v.fixed.assign(d->fixed, d->fixed + d->count);
    v.d.global_ids = nullptr;
    v.d.fixed = nullptr;
    const size_t n = (size_t)std::max<int64_t>(1, d->count) * v.d.parameters;
    GB_CUDA(ctx, cudaMalloc((void **)&v.params, n * sizeof(T)));
    GB_CUDA(ctx, cudaMalloc((void **)&v.backup, n * sizeof(T)));
    V.push_back(v);
    free_structure();
    return (int)V.size() - 1;
  }
  int add_factor_set(const gb_factor_set_desc *d, gb_graph_factor_fn fn, void *user) override {
    GB_TRY(require(d && d->residual_dim >= 1 && d->residual_dim <= GB_MAX_RESIDUAL && d->arity >= 1 && d->arity <= GB_MAX_ARITY &&
                       d->count >= 0 && d->count < (1ll << 31), "bad factor set"));
    GB_TRY(require((int)F.size() < MAX_SETS, "too many factor sets"));
    GB_TRY(require(fn != nullptr, "factor callback missing"));
    GB_TRY(require(d->vertex_index || d->count == 0, "vertex_index missing"));
    GB_TRY(require(d->loss == GB_LOSS_DEFAULT || (d->loss == GB_LOSS_HUBER && d->loss_delta > 0), "bad loss"));
    FSetH f;
    f.d = *d;
    for (int s = 0; s < d->arity; s++) GB_TRY(require(d->vertex_set[s] >= 0 && d->vertex_set[s] < (int)V.size(), "factor set names an unknown vertex set"));
    f.vidx.assign(d->vertex_index, d->vertex_index + d->count * d->arity);
    for (int64_t i = 0; i < d->count; i++)
      for (int s = 0; s < d->arity; s++) {
        const int32_t vi = f.vidx[i * d->arity + s];
        if (vi < 0 || vi >= V[d->vertex_set[s]].d.count) return ctx->fail(GB_ERR_INVALID, "factor %ld slot %d: vertex index %d out of range", (long)i, s, vi);
      }
    f.level.assign((size_t)d->count, 0);
    if (d->active) f.level.assign(d->active, d->active + d->count);
    f.d.vertex_index = nullptr;
    f.d.active = nullptr;
    f.fn = fn;
    f.user = user;
    // precision matrices default to the identity (factor.hpp:397-405)
    const int E = d->residual_dim;
    std::vector<S> Pm((size_t)std::max<int64_t>(1, d->count) * E * E);
    std::vector<T> Pt(Pm.size(), T(0));
    for (int64_t i = 0; i < d->count; i++)
      for (int a = 0; a < E; a++) Pt[(i * E + a) * E + a] = T(1);
    GB_CUDA(ctx, cudaMalloc((void **)&f.P, Pm.size() * sizeof(S)));
    F.push_back(f);
    const int id = (int)F.size() - 1;
    GB_TRY(upload_precision(id, Pt.data()));
    free_structure();
    return id;
  }
  int upload_precision(int id, const T *host) {
    FSetH &f = F[id];
    const size_t n = (size_t)f.d.count * f.d.residual_dim * f.d.residual_dim;
    if (n == 0) return GB_OK;
    if constexpr (std::is_same<T, S>::value) {
      GB_CUDA(ctx, cudaMemcpyAsync(f.P, host, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    } else {
      T *tmp = nullptr;
      GB_CUDA(ctx, cudaMalloc((void **)&tmp, n * sizeof(T)));
      cudaError_t e = cudaMemcpyAsync(tmp, host, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream);
      if (e == cudaSuccess) {
        k_g_from_T<T, S><<<grid_for((long long)n), 256, 0, ctx->stream>>>((long long)n, tmp, f.P);
        GB_LAUNCH(ctx);
        e = cudaStreamSynchronize(ctx->stream);
      }
      cudaFree(tmp);
      GB_CUDA(ctx, e);
    }
    GB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return GB_OK;
  }
  int set_update(int vs, gb_graph_update_fn fn, void *user) override {
    GB_TRY(require(vs >= 0 && vs < (int)V.size(), "unknown vertex set"));
    V[vs].upd = fn;
    V[vs].upd_user = user;
    return GB_OK;
  }
  int set_fixed(int vs, const uint8_t *fx) override {
    GB_TRY(require(vs >= 0 && vs < (int)V.size() && fx, "unknown vertex set"));
    V[vs].fixed.assign(fx, fx + V[vs].d.count);
    free_structure();
    return GB_OK;
  }
  int set_active(int fsi, const uint8_t *a) override {
    GB_TRY(require(fsi >= 0 && fsi < (int)F.size() && a, "unknown factor set"));
    F[fsi].level.assign(a, a + F[fsi].d.count);
    free_structure();
    return GB_OK;
  }
  int set_vertices(int vs, const void *h) override {
    GB_TRY(require(vs >= 0 && vs < (int)V.size() && h, "unknown vertex set"));
    VSetH &v = V[vs];
    GB_CUDA(ctx, cudaMemcpyAsync(v.params, h, (size_t)v.d.count * v.d.parameters * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    GB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    v.have_values = true;
    linearized = solved = false;
    return GB_OK;
  }
  int get_vertices(int vs, void *h) override {
    GB_TRY(require(vs >= 0 && vs < (int)V.size() && h, "unknown vertex set"));
    VSetH &v = V[vs];
    GB_CUDA(ctx, cudaMemcpyAsync(h, v.params, (size_t)v.d.count * v.d.parameters * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
    GB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return GB_OK;
  }
  int vertices_device(int vs, void **ptr) override {
    GB_TRY(require(vs >= 0 && vs < (int)V.size() && ptr, "unknown vertex set"));
    *ptr = V[vs].params;
    return GB_OK;
  }
  int set_precision(int fsi, const void *h) override {
    GB_TRY(require(fsi >= 0 && fsi < (int)F.size(), "unknown factor set"));
    FSetH &f = F[fsi];
    const int E = f.d.residual_dim;
    linearized = solved = false;
    if (h) return upload_precision(fsi, (const T *)h);
    std::vector<T> Pt((size_t)f.d.count * E * E, T(0));
    for (int64_t i = 0; i < f.d.count; i++)
      for (int a = 0; a < E; a++) Pt[(i * E + a) * E + a] = T(1);
    return upload_precision(fsi, Pt.data());
  }
  int set_loss(int fsi, int loss, double delta) override {
    GB_TRY(require(fsi >= 0 && fsi < (int)F.size(), "unknown factor set"));
    GB_TRY(require(loss == GB_LOSS_DEFAULT || (loss == GB_LOSS_HUBER && delta > 0), "bad loss"));
    F[fsi].d.loss = loss;
    F[fsi].d.loss_delta = delta;
    linearized = solved = false;
    if (initialized) GB_TRY(upload_tables());
    return GB_OK;
  }
  int set_scaling(int on) override {
    scale_on = on != 0;
    linearized = solved = false;
    return GB_OK;
  }

  DFSet<T, S> make_dfset(const FSetH &f) const {
    DFSet<T, S> d{};
    d.E = f.d.residual_dim;
    d.arity = f.d.arity;
    d.loss = f.d.loss;
    d.delta = f.d.loss_delta;
    for (int s = 0; s < f.d.arity; s++) {
      d.d[s] = V[f.d.vertex_set[s]].d.dimension;
      d.vset[s] = f.d.vertex_set[s];
      d.J[s] = f.bound_J[s] ? const_cast<S *>(f.bound_J[s]) : f.J[s];
    }
    d.count = f.d.count;
    d.nactive = (long long)f.active_idx.size();
    d.roff = f.roff;
    d.vidx = f.vidx_d;
    d.active_idx = f.active_idx_d;
    d.r = f.r;
    d.chi2 = f.chi2;
    d.dL = f.bound_dL ? const_cast<S *>(f.bound_dL) : f.dL;
    d.P = f.bound_P ? const_cast<S *>(f.bound_P) : f.P;
    return d;
  }
  DVSet<T> make_dvset(const VSetH &v) const {
    DVSet<T> d{};
    d.dim = v.d.dimension;
    d.npar = v.d.parameters;
    d.count = v.d.count;
    d.params = v.params;
    d.active = v.active_d;
    d.hoff = v.hoff_d;
    d.inc_ptr = v.inc_ptr_d;
    d.inc = v.inc_d;
    d.blockdiag = v.blockdiag;
    d.pinv = v.pinv;
    d.sdiag = v.sdiag;
    return d;
  }
  int upload_tables() {
    std::vector<DFSet<T, S>> ft;
    std::vector<DVSet<T>> vt;
    for (auto &f : F) ft.push_back(make_dfset(f));
    for (auto &v : V) vt.push_back(make_dvset(v));
    GB_CUDA(ctx, cudaMemcpyAsync(fs_d, ft.data(), ft.size() * sizeof(DFSet<T, S>), cudaMemcpyHostToDevice, ctx->stream));
    GB_CUDA(ctx, cudaMemcpyAsync(vs_d, vt.data(), vt.size() * sizeof(DVSet<T>), cudaMemcpyHostToDevice, ctx->stream));
    GB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return GB_OK;
  }

  // Graph::initialize_optimization + build_structure + Hessian::build_structure
  int initialize(int lvl, int64_t *info) override {
    GB_TRY(require(!V.empty() && !F.empty(), "graph needs at least one vertex set and one factor set"));
    GB_TRY(require(lvl >= 0 && lvl <= 127, "optimisation level out of range"));
    free_structure();
    bytes = 0;
    level = lvl;
    if (!h_pin) GB_CUDA(ctx, cudaMallocHost((void **)&h_pin, 64 * sizeof(double)));
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    // factor activity at this level (active.hpp:11-16); vertices referenced by an active factor (graph.hpp:171-210)
    std::vector<std::vector<uint8_t>> used(V.size());
    for (size_t i = 0; i < V.size(); i++) used[i].assign((size_t)V[i].d.count, 0);
    nactive_total = 0;
    rows = 0;
    for (auto &f : F) {
      f.active.assign((size_t)f.d.count, 0);
      f.active_idx.clear();
      for (int64_t i = 0; i < f.d.count; i++) {
        const uint8_t a = f.level[i];
        if ((a & 0x7F) <= lvl && (a & 0x80) == 0) {
          f.active[i] = 1;
          f.active_idx.push_back((int32_t)i);
          for (int s = 0; s < f.d.arity; s++) used[f.d.vertex_set[s]][f.vidx[i * f.d.arity + s]] = 1;
        }
      }
      f.roff = rows;
      rows += f.d.count * f.d.residual_dim;
      nactive_total += (long long)f.active_idx.size();
    }
    // block order: non-eliminated sets first, ascending global id (graph.hpp:112-147); active vertices get columns
    struct Entry { int64_t gid; int set; int64_t local; bool elim; };
    std::vector<Entry> order;
    for (size_t i = 0; i < V.size(); i++)
      for (int64_t k = 0; k < V[i].d.count; k++) order.push_back(Entry{V[i].gids[k], (int)i, k, V[i].d.eliminate != 0});
    std::sort(order.begin(), order.end(), [](const Entry &a, const Entry &b) {
      if (a.elim != b.elim) return !a.elim;
      return a.gid < b.gid;
    });
    for (size_t i = 1; i < order.size(); i++)
      if (order[i].gid == order[i - 1].gid && order[i].elim == order[i - 1].elim)
        return ctx->fail(GB_ERR_INVALID, "duplicate global vertex id %ld", (long)order[i].gid);
    for (size_t i = 0; i < V.size(); i++) {
      V[i].active.assign((size_t)V[i].d.count, 0);
      V[i].hoff.assign((size_t)V[i].d.count, -1);
      V[i].block.assign((size_t)V[i].d.count, -1);
    }
    dimH = 0;
    nblockcols = 0;
    std::vector<int> block_dim;
    std::vector<unsigned long long> colmap;
    for (const Entry &e : order) {
      VSetH &v = V[e.set];
      if (v.fixed[e.local] || !used[e.set][e.local]) continue;
      v.active[e.local] = 1;
      v.hoff[e.local] = dimH;
      v.block[e.local] = nblockcols;
      for (int c = 0; c < v.d.dimension; c++)
        colmap.push_back(((unsigned long long)e.set << 40) | ((unsigned long long)e.local << 8) | (unsigned long long)c);
      block_dim.push_back(v.d.dimension);
      dimH += v.d.dimension;
      nblockcols++;
    }
    GB_TRY(require(dimH > 0, "no active vertex: nothing to optimise"));
    GB_TRY(require(dimH < (1ll << 31), "Hessian dimension too large"));
    // incidence lists (vertex -> active factors touching it, in (factor set, factor, slot) order) and Hessian blocks
    std::vector<std::vector<long long>> inc_ptr(V.size());
    for (size_t i = 0; i < V.size(); i++) inc_ptr[i].assign((size_t)V[i].d.count + 1, 0);
    for (auto &f : F)
      for (int32_t fi : f.active_idx)
        for (int s = 0; s < f.d.arity; s++) {
          const int vs = f.d.vertex_set[s];
          const int32_t vi = f.vidx[(int64_t)fi * f.d.arity + s];
          if (V[vs].active[vi]) inc_ptr[vs][vi + 1]++;
        }
    std::vector<std::vector<unsigned long long>> inc(V.size());
    for (size_t i = 0; i < V.size(); i++) {
      for (int64_t k = 0; k < V[i].d.count; k++) inc_ptr[i][k + 1] += inc_ptr[i][k];
      inc[i].assign((size_t)inc_ptr[i][V[i].d.count], 0);
    }
    {
      std::vector<std::vector<long long>> cur = inc_ptr;
      for (size_t fs = 0; fs < F.size(); fs++) {
        auto &f = F[fs];
        for (int32_t fi : f.active_idx)
          for (int s = 0; s < f.d.arity; s++) {
            const int vs = f.d.vertex_set[s];
            const int32_t vi = f.vidx[(int64_t)fi * f.d.arity + s];
            if (V[vs].active[vi])
              inc[vs][cur[vs][vi]++] = ((unsigned long long)fs << 56) | ((unsigned long long)s << 48) | (unsigned long long)(uint32_t)fi;
          }
      }
    }
    // upper-triangular block coordinates: every active factor, every slot pair i <= j with both vertices active
    // (factor.hpp:657-694); sorted by (col, row), unique (hessian.hpp:48-85); offsets = prefix sums (hessian.hpp:270-278)
    struct Coord { int64_t col, row; int fset, si, sj; int64_t f; };
    std::vector<Coord> coords;
    for (size_t fs = 0; fs < F.size(); fs++) {
      auto &f = F[fs];
      for (int32_t fi : f.active_idx)
        for (int i = 0; i < f.d.arity; i++)
          for (int j = i; j < f.d.arity; j++) {
            const VSetH &vi = V[f.d.vertex_set[i]], &vj = V[f.d.vertex_set[j]];
            const int32_t a = f.vidx[(int64_t)fi * f.d.arity + i], b = f.vidx[(int64_t)fi * f.d.arity + j];
            if (!vi.active[a] || !vj.active[b]) continue;
            const int64_t bi = vi.block[a], bj = vj.block[b];
            if (bi > bj) coords.push_back(Coord{bi, bj, (int)fs, j, i, fi}); // transposed: the row vertex is slot j
            else coords.push_back(Coord{bj, bi, (int)fs, i, j, fi});
          }
    }
    std::stable_sort(coords.begin(), coords.end(), [](const Coord &a, const Coord &b) {
      if (a.col != b.col) return a.col < b.col;
      return a.row < b.row;
    });
    h_colptr.assign((size_t)nblockcols + 1, 0);
    h_rowidx.clear();
    h_offsets.clear();
    std::vector<long long> blk_ptr;
    std::vector<int> blk_rows, blk_cols;
    std::vector<HContrib> contrib;
    h_nvalues = 0;
    for (size_t i = 0; i < coords.size(); i++) {
      const bool fresh = i == 0 || coords[i].col != coords[i - 1].col || coords[i].row != coords[i - 1].row;
      if (fresh) {
        h_colptr[coords[i].col + 1]++;
        h_rowidx.push_back(coords[i].row);
        h_offsets.push_back(h_nvalues);
        blk_ptr.push_back((long long)contrib.size());
        blk_rows.push_back(block_dim[coords[i].row]);
        blk_cols.push_back(block_dim[coords[i].col]);
        h_nvalues += (long long)block_dim[coords[i].row] * block_dim[coords[i].col];
      }
      contrib.push_back(HContrib{coords[i].fset, coords[i].si, coords[i].sj, 0, coords[i].f});
    }
    blk_ptr.push_back((long long)contrib.size());
    for (int64_t c = 0; c < nblockcols; c++) h_colptr[c + 1] += h_colptr[c];
    std::vector<long long> scalar_block((size_t)h_nvalues);
    for (size_t bk = 0; bk < h_rowidx.size(); bk++) {
      const long long n = (long long)blk_rows[bk] * blk_cols[bk];
      for (long long k = 0; k < n; k++) scalar_block[h_offsets[bk] + k] = (long long)bk;
    }
    // device structure
    for (size_t i = 0; i < V.size(); i++) {
      VSetH &v = V[i];
      std::vector<int> ho(v.hoff.begin(), v.hoff.end());
      GB_TRY(upload(&v.active_d, v.active));
      GB_TRY(upload(&v.hoff_d, ho));
      GB_TRY(upload(&v.inc_ptr_d, inc_ptr[i]));
      GB_TRY(upload(&v.inc_d, inc[i]));
      const size_t nb = (size_t)v.d.count * v.d.dimension * v.d.dimension;
      GB_TRY(dalloc(&v.blockdiag, nb));
      GB_TRY(dalloc(&v.pinv, nb));
      GB_TRY(dalloc(&v.sdiag, (size_t)v.d.count * v.d.dimension));
      GB_TRY(dalloc(&v.delta, (size_t)v.d.count * v.d.dimension));
      GB_CUDA(ctx, cudaMemsetAsync(v.pinv, 0, std::max<size_t>(1, nb) * sizeof(IP), ctx->stream));
    }
    for (auto &f : F) {
      const int E = f.d.residual_dim;
      GB_TRY(upload(&f.vidx_d, f.vidx));
      GB_TRY(upload(&f.active_idx_d, f.active_idx));
      GB_TRY(upload(&f.active_d, f.active));
      GB_TRY(dalloc(&f.r, (size_t)f.d.count * E));
      GB_TRY(dalloc(&f.r_trial, (size_t)f.d.count * E));
      GB_TRY(dalloc(&f.chi2, (size_t)f.d.count));
      GB_TRY(dalloc(&f.dL, (size_t)f.d.count));
      GB_CUDA(ctx, cudaMemsetAsync(f.r, 0, std::max<size_t>(1, (size_t)f.d.count * E) * sizeof(T), ctx->stream));
      GB_CUDA(ctx, cudaMemsetAsync(f.r_trial, 0, std::max<size_t>(1, (size_t)f.d.count * E) * sizeof(T), ctx->stream));
      GB_CUDA(ctx, cudaMemsetAsync(f.chi2, 0, std::max<size_t>(1, (size_t)f.d.count) * sizeof(T), ctx->stream));
      GB_CUDA(ctx, cudaMemsetAsync(f.dL, 0, std::max<size_t>(1, (size_t)f.d.count) * sizeof(S), ctx->stream));
      for (int s = 0; s < f.d.arity; s++) {
        const size_t n = (size_t)f.d.count * E * V[f.d.vertex_set[s]].d.dimension;
        GB_TRY(dalloc(&f.J[s], n));
        if constexpr (std::is_same<T, S>::value) f.Jt[s] = (T *)f.J[s];
        else GB_TRY(dalloc(&f.Jt[s], n));
      }
    }
    GB_TRY(upload(&colmap_d, colmap));
    GB_TRY(upload(&blk_ptr_d, blk_ptr));
    GB_TRY(upload(&blk_off_d, std::vector<long long>(h_offsets.begin(), h_offsets.end())));
    GB_TRY(upload(&blk_rows_d, blk_rows));
    GB_TRY(upload(&blk_cols_d, blk_cols));
    GB_TRY(upload(&scalar_block_d, scalar_block));
    GB_TRY(upload(&contrib_d, contrib));
    GB_TRY(dalloc(&fs_d, F.size()));
    GB_TRY(dalloc(&vs_d, V.size()));
    T **vecs[] = {&b, &scales, &cdiag, &x, &xb, &r, &z, &p, &v2, &tmpH};
    for (T **vp : vecs) {
      GB_TRY(dalloc(vp, (size_t)dimH));
      GB_CUDA(ctx, cudaMemsetAsync(*vp, 0, (size_t)dimH * sizeof(T), ctx->stream));
    }
    GB_TRY(dalloc(&v1, (size_t)std::max<long long>(1, rows)));
    GB_CUDA(ctx, cudaMemsetAsync(v1, 0, (size_t)std::max<long long>(1, rows) * sizeof(T), ctx->stream));
    GB_TRY(dalloc(&partial, (size_t)3 * NPART * 8));
    GB_TRY(dalloc(&scal, 16));
    GB_TRY(dalloc(&state_d, 1));
    GB_TRY(upload_tables());
    int occ = 1;
    GB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_g_pcg<T, S, IP>, 256, 0));
    coop_blocks_per_sm = std::max(1, occ);
    initialized = true;
    if (info) {
      info[0] = dimH; info[1] = nblockcols; info[2] = (int64_t)h_rowidx.size(); info[3] = h_nvalues;
      info[4] = rows; info[5] = nactive_total; info[6] = bytes; info[7] = 0;
    }
    return GB_OK;
  }
  int vertex_columns(int vs, int64_t *out) override {
    GB_TRY(require(initialized && vs >= 0 && vs < (int)V.size() && out, "gb_graph_vertex_columns: not initialised / unknown set"));
    std::copy(V[vs].hoff.begin(), V[vs].hoff.end(), out);
    return GB_OK;
  }
  int hessian_structure(int64_t *cp, int64_t *ri, int64_t *off) override {
    GB_TRY(require(initialized, "graph is not initialised"));
    if (cp) std::copy(h_colptr.begin(), h_colptr.end(), cp);
    if (ri) std::copy(h_rowidx.begin(), h_rowidx.end(), ri);
    if (off) std::copy(h_offsets.begin(), h_offsets.end(), off);
    return GB_OK;
  }

  // run the caller's kernels for every factor set (Graph::compute_error / FactorDescriptor::compute_jacobians)
  int evaluate(bool with_jac, bool trial) {
    for (size_t i = 0; i < F.size(); i++) {
      FSetH &f = F[i];
      gb_graph_eval ev{};
      ev.factor_set = (int)i;
      ev.with_jacobians = with_jac ? 1 : 0;
      ev.num_factors = f.d.count;
      ev.num_active = (int64_t)f.active_idx.size();
      ev.active_index = f.active_idx_d;
      ev.vertex_index = f.vidx_d;
      for (int s = 0; s < f.d.arity; s++) {
        ev.vertices[s] = V[f.d.vertex_set[s]].params;
        ev.jacobians[s] = with_jac ? (void *)f.Jt[s] : nullptr;
      }
      ev.residuals = trial ? f.r_trial : f.r;
      ev.stream = (void *)ctx->stream;
      if (ev.num_active == 0) continue;
      const int rc = f.fn(&ev, f.user);
      if (rc != 0) return ctx->fail(GB_ERR_INVALID, "factor callback of set %d returned %d", (int)i, rc);
    }
    return GB_OK;
  }
  // chi2 of all active factors -> scal[slot]; linearisation point: also chi2_f and dL
  int enqueue_chi2(bool trial, int slot) {
    bool first = true;
    for (auto &f : F) {
      if (f.active_idx.empty()) continue;
      const int g = std::min(grid_for((long long)f.active_idx.size()), NPART);
      k_g_chi2<T, S><<<g, 256, 256 * sizeof(T), ctx->stream>>>(make_dfset(f), trial ? f.r_trial : f.r, trial ? nullptr : f.chi2,
                                                              trial ? nullptr : f.dL, partial);
      GB_TRY(launched());
      k_g_sum<T><<<1, 1, 0, ctx->stream>>>(partial, g, scal + slot, first ? 0 : 1);
      GB_TRY(launched());
      first = false;
    }
    return GB_OK;
  }
  int read_scalars(int n) {
    // scal is T; convert on the host
    std::vector<T> tmp(n);
    GB_CUDA(ctx, cudaMemcpyAsync(tmp.data(), scal, n * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
    GB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < n; i++) h_pin[i] = (double)tmp[i];
    return GB_OK;
  }
  int check_ready() {
    GB_TRY(require(initialized, "graph is not initialised (gb_graph_initialize)"));
    for (auto &v : V) GB_TRY(require(v.have_values || v.d.count == 0, "vertex values missing (gb_graph_set_vertices)"));
    return GB_OK;
  }

  // Graph::linearize (graph.hpp:236-290) + BlockJacobiPreconditioner::update_values (block_jacobi.hpp:79-112)
  int enqueue_linearize() {
    cudaStream_t st = ctx->stream;
    GB_TRY(evaluate(true, false));
    for (auto &f : F)
      for (int s = 0; s < f.d.arity; s++) {
        const long long n = f.d.count * f.d.residual_dim * V[f.d.vertex_set[s]].d.dimension;
        k_g_store_jac<T, S><<<grid_for(n), 256, 0, st>>>(make_dfset(f), s, f.Jt[s], f.active_d, V[f.d.vertex_set[s]].active_d);
        GB_TRY(launched());
      }
    GB_TRY(enqueue_chi2(false, 0));
    if (scale_on) {
      k_g_columns<T, S, 0><<<grid_for(dimH), 256, 0, st>>>(fs_d, vs_d, colmap_d, dimH, nullptr, scales, 1);
      GB_TRY(launched());
      for (auto &f : F)
        for (int s = 0; s < f.d.arity; s++) {
          if (f.active_idx.empty()) continue;
          k_g_scale_jac<T, S><<<grid_for((long long)f.active_idx.size() * V[f.d.vertex_set[s]].d.dimension), 256, 0, st>>>(
              make_dfset(f), s, scales, V[f.d.vertex_set[s]].hoff_d);
          GB_TRY(launched());
        }
    } else {
      std::vector<T> ones((size_t)dimH, T(1));
      GB_CUDA(ctx, cudaMemcpyAsync(scales, ones.data(), (size_t)dimH * sizeof(T), cudaMemcpyHostToDevice, st));
      GB_CUDA(ctx, cudaStreamSynchronize(st));
    }
    k_g_columns<T, S, 1><<<grid_for(dimH), 256, 0, st>>>(fs_d, vs_d, colmap_d, dimH, nullptr, b, 0);
    GB_TRY(launched());
    for (size_t i = 0; i < V.size(); i++) {
      const long long n = V[i].d.count * V[i].d.dimension * V[i].d.dimension;
      if (n == 0) continue;
      k_g_blockdiag<T, S, IP><<<grid_for(n), 256, 0, st>>>(fs_d, vs_d, (int)i);
      GB_TRY(launched());
    }
    // clamped scalar diagonal of the scaled system for the PCG damping term (pcg.hpp:93-104)
    k_g_columns<T, S, 0><<<grid_for(dimH), 256, 0, st>>>(fs_d, vs_d, colmap_d, dimH, nullptr, tmpH, 0);
    GB_TRY(launched());
    k_g_clamp<T><<<grid_for(dimH), 256, 0, st>>>(dimH, tmpH, cdiag, (T)1.0e-6, (T)1.0e32);
    GB_TRY(launched());
    linearized = true;
    damped = solved = false;
    return GB_OK;
  }
  int linearize(double *chi2) override {
    GB_TRY(require(!bound, "a caller-owned linearisation is bound: unbind it (gb_graph_bind_linearization with NULL) first"));
    GB_TRY(check_ready());
    GB_TRY(enqueue_linearize());
    GB_TRY(read_scalars(1));
    if (chi2) *chi2 = h_pin[0];
    return GB_OK;
  }
  int cost(double *chi2) override {
    GB_TRY(check_ready());
    GB_TRY(evaluate(false, true));
    GB_TRY(enqueue_chi2(true, 1));
    GB_TRY(read_scalars(2));
    if (chi2) *chi2 = h_pin[1];
    return GB_OK;
  }
  int d2h(void *dst, const void *src, size_t bytes_) {
    GB_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes_, cudaMemcpyDeviceToHost, ctx->stream));
    GB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return GB_OK;
  }
  template <typename X> int export_as_T(const X *src, long long n, void *out) {
    if constexpr (std::is_same<X, T>::value) return d2h(out, src, (size_t)n * sizeof(T));
    else {
      T *tmp = nullptr;
      GB_CUDA(ctx, cudaMalloc((void **)&tmp, (size_t)std::max<long long>(1, n) * sizeof(T)));
      k_g_to_T<T, X><<<grid_for(n), 256, 0, ctx->stream>>>(n, src, tmp);
      GB_LAUNCH(ctx);
      const int rc = d2h(out, tmp, (size_t)n * sizeof(T));
      cudaFree(tmp);
      return rc;
    }
  }
  int get(int which, int set, void *out) override {
    GB_TRY(require(initialized && linearized && out, "gb_graph_get needs a linearised graph"));
    if (which == 0) return d2h(out, b, (size_t)dimH * sizeof(T));
    if (which == 1) return d2h(out, scales, (size_t)dimH * sizeof(T));
    if (which == 17) return d2h(out, tmpH, (size_t)dimH * sizeof(T));
    if (which == 16) {
      GB_TRY(require(set >= 0 && set < (int)V.size(), "unknown vertex set"));
      if (damped) { // the stored blocks carry the damped diagonal; rebuild the undamped ones
        k_g_blockdiag<T, S, IP><<<grid_for(V[set].d.count * V[set].d.dimension * V[set].d.dimension), 256, 0, ctx->stream>>>(fs_d, vs_d, set);
        GB_TRY(launched());
      }
      return export_as_T<IP>(V[set].blockdiag, V[set].d.count * V[set].d.dimension * V[set].d.dimension, out);
    }
    GB_TRY(require(set >= 0 && set < (int)F.size(), "unknown factor set"));
    FSetH &f = F[set];
    if (which == 2) return d2h(out, f.r, (size_t)f.d.count * f.d.residual_dim * sizeof(T));
    if (which == 3) return d2h(out, f.chi2, (size_t)f.d.count * sizeof(T));
    if (which == 4) return export_as_T<S>(f.dL, f.d.count, out);
    if (which >= 5 && which < 5 + f.d.arity) {
      const int s = which - 5;
      return export_as_T<S>(f.J[s], f.d.count * f.d.residual_dim * V[f.d.vertex_set[s]].d.dimension, out);
    }
    return ctx->fail(GB_ERR_INVALID, "gb_graph_get: unknown selector %d", which);
  }
  int hessian_values(void *out) override {
    GB_TRY(require(initialized && linearized && out, "gb_graph_hessian_values needs a linearised graph"));
    S *H = nullptr;
    GB_CUDA(ctx, cudaMalloc((void **)&H, (size_t)std::max<long long>(1, h_nvalues) * sizeof(S)));
    k_g_hessian<T, S><<<grid_for(h_nvalues), 256, 0, ctx->stream>>>(fs_d, blk_ptr_d, contrib_d, blk_off_d, blk_rows_d, blk_cols_d,
                                                                     (long long)h_rowidx.size(), scalar_block_d, h_nvalues, H);
    GB_LAUNCH(ctx);
    int rc;
    if constexpr (std::is_same<S, __nv_bfloat16>::value) rc = export_as_T<S>(H, h_nvalues, out);
    else rc = d2h(out, H, (size_t)h_nvalues * sizeof(S));
    cudaFree(H);
    return rc;
  }
  int jv(const void *xh, void *yh) override {
    GB_TRY(require(initialized && linearized && xh && yh, "gb_graph_jv needs a linearised graph"));
    cudaStream_t st = ctx->stream;
    GB_CUDA(ctx, cudaMemcpyAsync(tmpH, xh, (size_t)dimH * sizeof(T), cudaMemcpyHostToDevice, st));
    GB_CUDA(ctx, cudaMemsetAsync(v1, 0, (size_t)std::max<long long>(1, rows) * sizeof(T), st));
    k_g_jv<T, S><<<grid_for(rows), 256, 0, st>>>(fs_d, (int)F.size(), vs_d, tmpH, v1);
    GB_TRY(launched());
    GB_TRY(d2h(yh, v1, (size_t)rows * sizeof(T)));
    return refresh_tmpH();
  }
  int jtpv(const void *vh, void *yh) override {
    GB_TRY(require(initialized && linearized && vh && yh, "gb_graph_jtpv needs a linearised graph"));
    cudaStream_t st = ctx->stream;
    GB_CUDA(ctx, cudaMemcpyAsync(v1, vh, (size_t)rows * sizeof(T), cudaMemcpyHostToDevice, st));
    k_g_columns<T, S, 2><<<grid_for(dimH), 256, 0, st>>>(fs_d, vs_d, colmap_d, dimH, v1, v2, 0);
    GB_TRY(launched());
    return d2h(yh, v2, (size_t)dimH * sizeof(T));
  }
  int refresh_tmpH() { // tmpH doubles as the export of the scaled scalar diagonal (selector 17)
    k_g_columns<T, S, 0><<<grid_for(dimH), 256, 0, ctx->stream>>>(fs_d, vs_d, colmap_d, dimH, nullptr, tmpH, 0);
    return launched();
  }

  int set_damping(double m, int ident) override {
    GB_TRY(require(std::isfinite(m) && m >= 0, "damping must be finite and non-negative"));
    mu = m;
    use_identity = ident;
    damped = false;
    return GB_OK;
  }
  // BlockJacobiPreconditioner::set_damping_factor (block_jacobi.hpp:114-168)
  int enqueue_damping() {
    for (size_t i = 0; i < V.size(); i++) {
      if (V[i].d.count == 0) continue;
      k_g_invert<T, IP><<<grid_for(V[i].d.count), 64, 0, ctx->stream>>>(vs_d, (int)i, mu, use_identity);
      GB_TRY(launched());
    }
    damped = true;
    return GB_OK;
  }
  int enqueue_pcg(const gb_pcg_options *o) {
    PcgArgs<T> A{};
    A.dimH = dimH; A.rows = rows; A.nf = (int)F.size(); A.nv = (int)V.size();
    A.max_iter = (int)o->max_iterations; A.use_identity = use_identity;
    A.mu = (T)mu; A.tol = (T)o->tolerance; A.ratio = (T)o->rejection_ratio;
    A.b = b_bound ? b_bound : b; A.cdiag = cdiag; A.x = x; A.xb = xb; A.r = r; A.z = z; A.p = p; A.v1 = v1; A.v2 = v2;
    A.partial = partial; A.colmap = colmap_d; A.state = state_d;
    const long long work = std::max<long long>(dimH, nactive_total * GB_MAX_RESIDUAL / 2);
    int blocks = (int)std::min<long long>((work + 255) / 256, (long long)nsm * coop_blocks_per_sm);
    blocks = std::max(1, std::min(blocks, NPART));
    const DFSet<T, S> *fsp = fs_d;
    const DVSet<T> *vsp = vs_d;
    void *args[] = {(void *)&fsp, (void *)&vsp, (void *)&A};
    GB_CUDA(ctx, cudaLaunchCooperativeKernel((void *)k_g_pcg<T, S, IP>, dim3(blocks), dim3(256), args, 0, ctx->stream));
    GB_LAUNCH(ctx);
    solved = true;
    return GB_OK;
  }
  int check_pcg(const gb_pcg_options *o) {
    GB_TRY(require(o && o->max_iterations >= 0 && o->max_iterations < (1 << 20) && o->tolerance >= 0 && o->rejection_ratio > 0, "bad PCG options"));
    return GB_OK;
  }
  int solve(const gb_pcg_options *o, void *delta_host, gb_solve_info *info) override {
    GB_TRY(require(initialized && linearized, "gb_graph_solve needs a linearised graph"));
    GB_TRY(check_pcg(o));
    GB_TRY(enqueue_damping());
    GB_TRY(enqueue_pcg(o));
    PcgState stt{};
    GB_TRY(d2h(&stt, state_d, sizeof(PcgState)));
    if (delta_host) GB_TRY(d2h(delta_host, x, (size_t)dimH * sizeof(T)));
    if (info) {
      info->pcg_iterations = stt.iterations;
      info->rz_final = stt.rz_final;
      info->stop_reason = stt.stop_reason;
      info->schur_mode = 0;
    }
    return GB_OK;
  }

  // backup_parameters + apply_update (graph.hpp:292-309)
  int enqueue_step() {
    cudaStream_t st = ctx->stream;
    for (size_t i = 0; i < V.size(); i++) {
      VSetH &v = V[i];
      if (v.d.count == 0) continue;
      GB_CUDA(ctx, cudaMemcpyAsync(v.backup, v.params, (size_t)v.d.count * v.d.parameters * sizeof(T), cudaMemcpyDeviceToDevice, st));
      k_g_step<T><<<grid_for(v.d.count * v.d.dimension), 256, 0, st>>>(make_dvset(v), x, scales, v.delta, v.upd ? 0 : 1);
      GB_TRY(launched());
      if (v.upd) {
        gb_graph_update u{};
        u.vertex_set = (int)i;
        u.count = v.d.count;
        u.vertices = v.params;
        u.delta = v.delta;
        u.active = v.active_d;
        u.stream = (void *)st;
        const int rc = v.upd(&u, v.upd_user);
        if (rc != 0) return ctx->fail(GB_ERR_INVALID, "update callback of vertex set %d returned %d", (int)i, rc);
      }
    }
    return GB_OK;
  }
  int enqueue_revert() {
    for (auto &v : V)
      if (v.d.count)
        GB_CUDA(ctx, cudaMemcpyAsync(v.params, v.backup, (size_t)v.d.count * v.d.parameters * sizeof(T), cudaMemcpyDeviceToDevice, ctx->stream));
    return GB_OK;
  }

  // ---- Solver<T,S> plug-in mode: the CALLER linearises (Graphite's own Graph::linearize through the user's traits); its
  //      device buffers are used in place: stored (scaled) Jacobians, loss derivatives, precision matrices, gradient ----
  int bind_linearization(int fsi, const void *const *jac, const void *dl, const void *prec) override {
    GB_TRY(require(initialized, "graph is not initialised (gb_graph_initialize)"));
    GB_TRY(require(fsi >= 0 && fsi < (int)F.size(), "unknown factor set"));
    FSetH &f = F[fsi];
    if (!jac) { // unbind
      for (int s = 0; s < GB_MAX_ARITY; s++) f.bound_J[s] = nullptr;
      f.bound_dL = f.bound_P = nullptr;
    } else {
      GB_TRY(require(dl != nullptr, "loss derivatives missing"));
      for (int s = 0; s < f.d.arity; s++) {
        GB_TRY(require(jac[s] != nullptr, "Jacobian pointer missing"));
        f.bound_J[s] = (const S *)jac[s];
      }
      f.bound_dL = (const S *)dl;
      f.bound_P = (const S *)prec; // null: the set's own precision matrices
    }
    bound = false;
    for (auto &g : F) bound = bound || g.bound_dL != nullptr;
    linearized = damped = solved = false;
    return upload_tables();
  }
  int bind_gradient(const void *bd) override {
    b_bound = (const T *)bd;
    solved = false;
    return GB_OK;
  }
  // Solver::update_values: block diagonals of the preconditioner and the clamped scalar diagonal from the bound Jacobians
  int update_values() override {
    GB_TRY(require(initialized && bound, "gb_graph_update_values needs a bound linearisation"));
    for (auto &f : F) GB_TRY(require(f.bound_dL != nullptr || f.active_idx.empty(), "every active factor set needs a bound linearisation"));
    GB_TRY(require(b_bound != nullptr, "gradient missing (gb_graph_bind_gradient)"));
    cudaStream_t st = ctx->stream;
    for (size_t i = 0; i < V.size(); i++) {
      const long long n = V[i].d.count * V[i].d.dimension * V[i].d.dimension;
      if (n == 0) continue;
      k_g_blockdiag<T, S, IP><<<grid_for(n), 256, 0, st>>>(fs_d, vs_d, (int)i);
      GB_TRY(launched());
    }
    k_g_columns<T, S, 0><<<grid_for(dimH), 256, 0, st>>>(fs_d, vs_d, colmap_d, dimH, nullptr, tmpH, 0);
    GB_TRY(launched());
    k_g_clamp<T><<<grid_for(dimH), 256, 0, st>>>(dimH, tmpH, cdiag, (T)1.0e-6, (T)1.0e32);
    GB_TRY(launched());
    linearized = true;
    damped = solved = false;
    return GB_OK;
  }
  int solve_device(const gb_pcg_options *o, void *delta_dev, gb_solve_info *info) override {
    GB_TRY(require(initialized && linearized && delta_dev, "gb_graph_solve_device needs a linearised graph"));
    GB_TRY(check_pcg(o));
    GB_TRY(enqueue_damping());
    GB_TRY(enqueue_pcg(o));
    GB_CUDA(ctx, cudaMemcpyAsync(delta_dev, x, (size_t)dimH * sizeof(T), cudaMemcpyDeviceToDevice, ctx->stream));
    PcgState stt{};
    GB_TRY(d2h(&stt, state_d, sizeof(PcgState)));
    if (info) {
      info->pcg_iterations = stt.iterations;
      info->rz_final = stt.rz_final;
      info->stop_reason = stt.stop_reason;
      info->schur_mode = 0;
    }
    return GB_OK;
  }

  int lm(const gb_lm_options *o, gb_lm_result *res_out, double *traj) override {
    GB_TRY(require(!bound, "a caller-owned linearisation is bound: the caller drives the LM loop"));
    GB_TRY(check_ready());
    GB_TRY(require(o && o->iterations >= 0, "bad LM options"));
    GB_TRY(check_pcg(&o->pcg));
    cudaStream_t st = ctx->stream;
    gb_lm_result R{};
    cudaEvent_t e0, e1;
    GB_CUDA(ctx, cudaEventCreate(&e0));
    GB_CUDA(ctx, cudaEventCreate(&e1));
    cudaEventRecord(e0, st);
    T mu_l = (T)o->initial_damping, nu = o->initial_nu > 0 ? (T)o->initial_nu : T(2);
    GB_TRY(enqueue_linearize());
    GB_TRY(read_scalars(1));
    T chi2 = (T)h_pin[0];
    R.initial_chi2 = (double)chi2;
    bool run = true;
    int num_bad = 0;
    int64_t it = 0;
    for (; it < o->iterations && run; it++) {
      mu = (double)mu_l;
      use_identity = o->use_identity;
      GB_TRY(enqueue_damping());
      GB_TRY(enqueue_pcg(&o->pcg));
      GB_TRY(enqueue_step());
      GB_TRY(evaluate(false, true));
      GB_TRY(enqueue_chi2(true, 1));
      const int g = std::min(grid_for(dimH), NPART);
      k_g_rho<T><<<g, 256, 256 * sizeof(T), st>>>(dimH, x, b, mu_l, partial);
      GB_TRY(launched());
      k_g_sum<T><<<1, 1, 0, st>>>(partial, g, scal + 2, 0);
      GB_TRY(launched());
      GB_TRY(read_scalars(3));
      PcgState stt{};
      GB_TRY(d2h(&stt, state_d, sizeof(PcgState)));
      R.pcg_iterations_total += stt.iterations;
      T new_chi2 = (T)h_pin[1];
      const T denom = (T)h_pin[2] + (T)1.0e-3;
      const T rho = (chi2 - new_chi2) / denom;
      const bool accepted_now = std::isfinite((double)new_chi2) && rho > T(0);
      if (accepted_now) {
        double alpha = 1.0 - std::pow(2.0 * (double)rho - 1.0, 3);
        alpha = std::max(std::min(alpha, 2.0 / 3.0), 1.0 / 3.0);
        mu_l *= (T)alpha;
        nu = T(2);
        GB_TRY(enqueue_linearize());
        R.accepted++;
      } else {
        GB_TRY(enqueue_revert());
        mu_l *= nu;
        nu *= T(2);
        new_chi2 = chi2;
        R.rejected++;
        solved = false;
      }
      if (traj) {
        traj[4 * it + 0] = (double)chi2; traj[4 * it + 1] = (double)new_chi2;
        traj[4 * it + 2] = (double)mu_l; traj[4 * it + 3] = (double)stt.iterations;
      }
      if (o->verbose) printf("%6ld %22.12g %22.12g %16.8g  pcg %ld\n", (long)it, (double)chi2, (double)new_chi2, (double)mu_l, (long)stt.iterations);
      const T initial_chi2 = chi2;
      chi2 = new_chi2;
      if (!std::isfinite((double)mu_l)) { run = false; R.termination = GB_LM_DAMPING_NOT_FINITE; }
      if (rho == T(0)) { it++; R.termination = GB_LM_RHO_ZERO; break; }
      if (o->stop_flag && *o->stop_flag) { it++; R.termination = GB_LM_STOP_FLAG; break; }
      if (o->early_stop && accepted_now) {
        if ((initial_chi2 - new_chi2) * T(1.0e3) < initial_chi2) num_bad++;
        else num_bad = 0;
        if (num_bad >= 3) { it++; R.termination = GB_LM_EARLY_STOP; break; }
      }
    }
    cudaEventRecord(e1, st);
    GB_CUDA(ctx, cudaStreamSynchronize(st));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    R.iterations = it;
    R.final_chi2 = (double)chi2;
    R.final_damping = (double)mu_l;
    R.final_nu = (double)nu;
    R.seconds_total = ms * 1e-3;
    if (res_out) *res_out = R;
    return GB_OK;
  }
};

} // namespace gg

struct gb_graph {
  gb_context *ctx;
  gg::GraphBase *impl;
};

#define GG(g)                                  \
  if (!(g) || !(g)->impl) return GB_ERR_INVALID; \
  cudaSetDevice((g)->ctx->device)

extern "C" {

int gb_graph_create(gb_context *ctx, int pT, int pS, gb_graph **out) {
  if (!ctx || !out) return GB_ERR_INVALID;
  *out = nullptr;
  cudaSetDevice(ctx->device);
  gg::GraphBase *impl = nullptr;
  if (pT == GB_F64 && pS == GB_F64) impl = new gg::Graph<double, double>();
  else if (pT == GB_F32 && pS == GB_F32) impl = new gg::Graph<float, float>();
  else if (pT == GB_F64 && pS == GB_F32) impl = new gg::Graph<double, float>();
  else if (pT == GB_F64 && pS == GB_BF16) impl = new gg::Graph<double, __nv_bfloat16>();
  else return ctx->fail(GB_ERR_UNSUPPORTED, "precision combination (T=%d, S=%d) is not supported", pT, pS);
  impl->ctx = ctx;
  *out = new gb_graph{ctx, impl};
  return GB_OK;
}
int gb_graph_destroy(gb_graph *g) {
  if (!g) return GB_ERR_INVALID;
  cudaSetDevice(g->ctx->device);
  cudaStreamSynchronize(g->ctx->stream);
  delete g->impl;
  delete g;
  return GB_OK;
}
int gb_graph_add_vertex_set(gb_graph *g, const gb_vertex_set_desc *d) { GG(g); return g->impl->add_vertex_set(d); }
int gb_graph_add_factor_set(gb_graph *g, const gb_factor_set_desc *d, gb_graph_factor_fn fn, void *u) { GG(g); return g->impl->add_factor_set(d, fn, u); }
int gb_graph_set_update(gb_graph *g, int vs, gb_graph_update_fn fn, void *u) { GG(g); return g->impl->set_update(vs, fn, u); }
int gb_graph_set_fixed(gb_graph *g, int vs, const uint8_t *f) { GG(g); return g->impl->set_fixed(vs, f); }
int gb_graph_set_active(gb_graph *g, int fs, const uint8_t *a) { GG(g); return g->impl->set_active(fs, a); }
int gb_graph_set_vertices(gb_graph *g, int vs, const void *h) { GG(g); return g->impl->set_vertices(vs, h); }
int gb_graph_get_vertices(gb_graph *g, int vs, void *h) { GG(g); return g->impl->get_vertices(vs, h); }
int gb_graph_vertices_device(gb_graph *g, int vs, void **p) { GG(g); return g->impl->vertices_device(vs, p); }
int gb_graph_set_precision(gb_graph *g, int fs, const void *h) { GG(g); return g->impl->set_precision(fs, h); }
int gb_graph_set_loss(gb_graph *g, int fs, int loss, double delta) { GG(g); return g->impl->set_loss(fs, loss, delta); }
int gb_graph_set_scaling(gb_graph *g, int on) { GG(g); return g->impl->set_scaling(on); }
int gb_graph_initialize(gb_graph *g, int level, int64_t info[8]) { GG(g); return g->impl->initialize(level, info); }
int gb_graph_vertex_columns(gb_graph *g, int vs, int64_t *c) { GG(g); return g->impl->vertex_columns(vs, c); }
int gb_graph_hessian_structure(gb_graph *g, int64_t *cp, int64_t *ri, int64_t *off) { GG(g); return g->impl->hessian_structure(cp, ri, off); }
int gb_graph_linearize(gb_graph *g, double *chi2) { GG(g); return g->impl->linearize(chi2); }
int gb_graph_cost(gb_graph *g, double *chi2) { GG(g); return g->impl->cost(chi2); }
int gb_graph_get(gb_graph *g, int which, int set, void *out) { GG(g); return g->impl->get(which, set, out); }
int gb_graph_hessian_values(gb_graph *g, void *v) { GG(g); return g->impl->hessian_values(v); }
int gb_graph_jv(gb_graph *g, const void *x, void *y) { GG(g); return g->impl->jv(x, y); }
int gb_graph_jtpv(gb_graph *g, const void *v, void *y) { GG(g); return g->impl->jtpv(v, y); }
int gb_graph_set_damping(gb_graph *g, double mu, int ident) { GG(g); return g->impl->set_damping(mu, ident); }
int gb_graph_solve(gb_graph *g, const gb_pcg_options *o, void *d, gb_solve_info *i) { GG(g); return g->impl->solve(o, d, i); }
int gb_graph_lm(gb_graph *g, const gb_lm_options *o, gb_lm_result *r, double *t) { GG(g); return g->impl->lm(o, r, t); }
int gb_graph_bind_linearization(gb_graph *g, int fs, const void *const *jac, const void *dl, const void *prec) {
  GG(g);
  return g->impl->bind_linearization(fs, jac, dl, prec);
}
int gb_graph_bind_gradient(gb_graph *g, const void *b) { GG(g); return g->impl->bind_gradient(b); }
int gb_graph_update_values(gb_graph *g) { GG(g); return g->impl->update_values(); }
int gb_graph_solve_device(gb_graph *g, const gb_pcg_options *o, void *d, gb_solve_info *i) { GG(g); return g->impl->solve_device(o, d, i); }

} // extern "C"

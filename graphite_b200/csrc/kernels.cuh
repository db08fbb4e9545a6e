// kernels.cuh — the sm_100a kernels of the LM inner loop.
//
// Layout (HBM).  Observations are sorted by (point, camera) and cut into TILES of whole points with at
// most TILE (=256) observations; one CTA owns one tile, one thread one observation.  Per observation:
//   Jc   9 planes of vec2<S>   (plane j = column j of the 2x9 camera Jacobian)   -> fully coalesced
//   Jp   3 planes of vec2<S>
//   r    vec2<T>
// Per point:  Cg[9] = {C00,C01,C02,C11,C12,C22, g0,g1,g2}  (C = sum Jp^T Jp, g = -sum Jp^T r, unscaled),
//             W[6] (= D_p (D_p C D_p + damping)^-1 D_p), h[3] = W g.
// Per camera: 10-padded rows (80 B in FP64) so that a gather is five 16-byte loads.
//
// Reductions are atomic-free and deterministic:
//   by point  - a point's observations are contiguous inside one tile: staged in shared memory,
//               summed sequentially by one thread per point;
//   by camera - each tile knows (structure build) the rank of every observation in the tile's
//               (camera, observation) order and the camera SEGMENTS of that order.  Threads stage their
//               9-vectors in shared memory at their rank, one thread per (segment, component) sums the
//               segment, the partial goes to part[segment]; a second kernel sums each camera's partials
//               in tile order (camera -> segment CSR).
// All arithmetic on the Jacobians is done in the UNSCALED space; the Jacobi scaling of the reference
// (graph.hpp:254-281) is applied as D_c / D_p on the camera- and point-sized quantities, which is the
// same algebra: J~ = J D  =>  J~^T J~ = D J^T J D.
#pragma once
#include <cfloat>
#include <cstdint>
#include <cuda_runtime.h>

#include "bal_math.cuh"

namespace gb {

constexpr int TILE = 256;
constexpr int CAM_STRIDE = 10; // padded camera row

template <typename T> struct V2;
template <> struct V2<double> {
  using type = double2;
  static __device__ __forceinline__ type make(double a, double b) { return make_double2(a, b); }
};
template <> struct V2<float> {
  using type = float2;
  static __device__ __forceinline__ type make(float a, float b) { return make_float2(a, b); }
};

struct TileStruct {
  int64_t M, Mpad;
  int32_t Nc, Np, ntiles, nseg;
  const int32_t *cam_idx, *pt_idx; // [M]
  const uint8_t *rank;             // [M] rank of the observation in its tile's (camera, obs) order
  const int32_t *pptr;             // [Np+1]
  const int32_t *tile_obs, *tile_pt, *tile_seg; // [ntiles+1]
  const int32_t *seg_cam;          // [nseg]
  const int32_t *seg_begin;        // [nseg+1] global sorted position where the segment starts
  const int32_t *cam_seg_ptr;      // [Nc+1]
  const int32_t *cam_seg_list;     // [nseg]
};

// ---------------------------------------------------------------------------------------------
// reductions
// ---------------------------------------------------------------------------------------------
// Deterministic block sum (fixed tree); result valid in every thread.  blockDim.x <= 1024.
template <typename T> __device__ __forceinline__ T block_sum(T v, T *sh /*[32]*/) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  T tot = T(0);
  for (int i = 0; i < nw; i++) tot += sh[i];
  return tot;
}

// Stage v[9] at row `rank`, then sum every camera segment of this tile; part row stride = pstride.
template <typename T>
__device__ __forceinline__ void tile_cam_reduce(const T v[9], bool active, int rank, int o0, int sg, int nsg,
                                                const int32_t *__restrict__ seg_begin, T *sv /*[TILE*9]*/,
                                                T *__restrict__ part, int pstride, int poff) {
  if (active) {
#pragma unroll
    for (int k = 0; k < 9; k++) sv[rank * 9 + k] = v[k];
  }
  __syncthreads();
  for (int item = threadIdx.x; item < nsg * 9; item += blockDim.x) {
    const int s = item / 9, k = item - 9 * s;
    const int b = seg_begin[sg + s] - o0, e = seg_begin[sg + s + 1] - o0;
    T acc = T(0);
    for (int row = b; row < e; row++) acc += sv[row * 9 + k];
    part[(int64_t)(sg + s) * pstride + poff + k] = acc;
  }
  __syncthreads();
}

// Sum the partials of one camera (block per camera, 288 threads = 32 sub-lists x 9 components).
// out[g*9+k] for g < ngroups; deterministic: sub-list i takes list entries i, i+32, ...; then 0..31 in order.
template <typename T>
__device__ __forceinline__ void cam_gather(const TileStruct &ts, int c, const T *__restrict__ part, int pstride,
                                           int ngroups, T *sh /*[32*9]*/, T *out /*smem [ngroups*9]*/) {
  const int sub = threadIdx.x / 9, k = threadIdx.x - 9 * sub;
  const int b = ts.cam_seg_ptr[c], e = ts.cam_seg_ptr[c + 1];
  for (int g = 0; g < ngroups; g++) {
    T acc = T(0);
    if (sub < 32)
      for (int i = b + sub; i < e; i += 32) acc += part[(int64_t)ts.cam_seg_list[i] * pstride + g * 9 + k];
    __syncthreads();
    if (sub < 32) sh[sub * 9 + k] = acc;
    __syncthreads();
    if (threadIdx.x < 9) {
      T tot = T(0);
      for (int i = 0; i < 32; i++) tot += sh[i * 9 + threadIdx.x];
      out[g * 9 + threadIdx.x] = tot;
    }
  }
  __syncthreads();
}

template <typename T> __device__ __forceinline__ void load_cam(const T *__restrict__ cams, int c, T *cam);
template <> __device__ __forceinline__ void load_cam<double>(const double *__restrict__ cams, int c, double *cam) {
  const double2 *p = reinterpret_cast<const double2 *>(cams + (int64_t)c * CAM_STRIDE);
#pragma unroll
  for (int i = 0; i < 5; i++) {
    const double2 v = __ldg(p + i);
    cam[2 * i] = v.x;
    cam[2 * i + 1] = v.y;
  }
}
template <> __device__ __forceinline__ void load_cam<float>(const float *__restrict__ cams, int c, float *cam) {
  const float2 *p = reinterpret_cast<const float2 *>(cams + (int64_t)c * CAM_STRIDE);
#pragma unroll
  for (int i = 0; i < 5; i++) {
    const float2 v = __ldg(p + i);
    cam[2 * i] = v.x;
    cam[2 * i + 1] = v.y;
  }
}

// ---------------------------------------------------------------------------------------------
// K1: factor evaluation + point-side assembly + camera-side partials of diag(B) and g_c
//     replaces compute_error_kernel / compute_jacobian_kernel / compute_chi2_kernel /
//     compute_hessian_scalar_diagonal_kernel / compute_b_kernel (ops/error.hpp:252, ops/linearize.hpp:10,238,
//     ops/chi2.hpp:32, ops/hessian.hpp:418)
// ---------------------------------------------------------------------------------------------
template <typename T, typename S>
__global__ void __launch_bounds__(TILE)
k_linearize(TileStruct ts, const T *__restrict__ cams, const T *__restrict__ pts,
            const typename V2<T>::type *__restrict__ obs, typename V2<S>::type *__restrict__ Jc,
            typename V2<S>::type *__restrict__ Jp, typename V2<T>::type *__restrict__ res, T *__restrict__ Cg,
            T *__restrict__ part /*[nseg][18]*/, double *__restrict__ cost_part) {
  __shared__ T sv[TILE * 9];
  __shared__ double shd[32];
  const int tile = blockIdx.x, t = threadIdx.x;
  const int o0 = ts.tile_obs[tile], n = ts.tile_obs[tile + 1] - o0;
  const int64_t o = (int64_t)o0 + t;
  const bool active = t < n;
  BalObs<T> B;
  int rank = 0;
  double cost = 0.0;
  if (active) {
    const int c = ts.cam_idx[o], p = ts.pt_idx[o];
    rank = ts.rank[o];
    T cam[10], X[3], ob[2];
    load_cam<T>(cams, c, cam);
    X[0] = pts[3 * (int64_t)p];
    X[1] = pts[3 * (int64_t)p + 1];
    X[2] = pts[3 * (int64_t)p + 2];
    const typename V2<T>::type ov = obs[o];
    ob[0] = ov.x;
    ob[1] = ov.y;
    bal_residual_jacobian<T>(cam, X, ob, B);
#pragma unroll
    for (int j = 0; j < 9; j++) Jc[(int64_t)j * ts.Mpad + o] = V2<S>::make((S)B.Jc[2 * j], (S)B.Jc[2 * j + 1]);
#pragma unroll
    for (int j = 0; j < 3; j++) Jp[(int64_t)j * ts.Mpad + o] = V2<S>::make((S)B.Jp[2 * j], (S)B.Jp[2 * j + 1]);
    res[o] = V2<T>::make(B.r[0], B.r[1]);
    cost = (double)(B.r[0] * B.r[0] + B.r[1] * B.r[1]);
    // what is stored is what every later kernel reads: keep the assembly consistent with S
#pragma unroll
    for (int j = 0; j < 18; j++) B.Jc[j] = (T)(S)B.Jc[j];
#pragma unroll
    for (int j = 0; j < 6; j++) B.Jp[j] = (T)(S)B.Jp[j];
    // point side: C (6 unique) and g = -Jp^T r
    T *row = sv + t * 9;
    row[0] = B.Jp[0] * B.Jp[0] + B.Jp[1] * B.Jp[1];
    row[1] = B.Jp[0] * B.Jp[2] + B.Jp[1] * B.Jp[3];
    row[2] = B.Jp[0] * B.Jp[4] + B.Jp[1] * B.Jp[5];
    row[3] = B.Jp[2] * B.Jp[2] + B.Jp[3] * B.Jp[3];
    row[4] = B.Jp[2] * B.Jp[4] + B.Jp[3] * B.Jp[5];
    row[5] = B.Jp[4] * B.Jp[4] + B.Jp[5] * B.Jp[5];
    row[6] = -(B.Jp[0] * B.r[0] + B.Jp[1] * B.r[1]);
    row[7] = -(B.Jp[2] * B.r[0] + B.Jp[3] * B.r[1]);
    row[8] = -(B.Jp[4] * B.r[0] + B.Jp[5] * B.r[1]);
  }
  __syncthreads();
  {
    const int p0 = ts.tile_pt[tile], npt = ts.tile_pt[tile + 1] - p0;
    // 9 threads per point: thread (q, k) sums component k of point q sequentially over its observations
    for (int item = t; item < npt * 9; item += TILE) {
      const int q = item / 9, k = item - 9 * q;
      const int b = ts.pptr[p0 + q] - o0, e = ts.pptr[p0 + q + 1] - o0;
      T acc = T(0);
      for (int rowi = b; rowi < e; rowi++) acc += sv[rowi * 9 + k];
      Cg[(int64_t)(p0 + q) * 9 + k] = acc;
    }
  }
  __syncthreads();
  const int sg = ts.tile_seg[tile], nsg = ts.tile_seg[tile + 1] - sg;
  T v[9];
#pragma unroll
  for (int k = 0; k < 9; k++) v[k] = active ? B.Jc[2 * k] * B.Jc[2 * k] + B.Jc[2 * k + 1] * B.Jc[2 * k + 1] : T(0);
  tile_cam_reduce<T>(v, active, rank, o0, sg, nsg, ts.seg_begin, sv, part, 18, 0);
#pragma unroll
  for (int k = 0; k < 9; k++) v[k] = active ? -(B.Jc[2 * k] * B.r[0] + B.Jc[2 * k + 1] * B.r[1]) : T(0);
  tile_cam_reduce<T>(v, active, rank, o0, sg, nsg, ts.seg_begin, sv, part, 18, 9);
  const double tot = block_sum<double>(cost, shd);
  if (t == 0) cost_part[tile] = tot;
}

// Camera side of linearize: diag(B), g_c, Jacobi scales s = 1/(eps + sqrt(diag)) (graph.hpp:262-270), b_c = s g_c.
template <typename T>
__global__ void __launch_bounds__(288)
k_cam_reduce_lin(TileStruct ts, const T *__restrict__ part, T *__restrict__ diagB, T *__restrict__ gc,
                 int do_finish, int scale_on, T *__restrict__ scale_c, T *__restrict__ b_c) {
  __shared__ T sh[32 * 9];
  __shared__ T out[18];
  const int c = blockIdx.x;
  cam_gather<T>(ts, c, part, 18, 2, sh, out);
  if (threadIdx.x < 9) {
    const int k = threadIdx.x;
    diagB[c * 9 + k] = out[k];
    gc[c * 9 + k] = out[9 + k];
    if (do_finish) {
      const T s = scale_on ? (T)(1.0 / (DBL_EPSILON + sqrt((double)out[k]))) : T(1);
      scale_c[c * 9 + k] = s;
      b_c[c * 9 + k] = s * out[9 + k];
    }
  }
}
// After a multi-GPU allreduce of diagB / gc.
template <typename T>
__global__ void k_cam_finish_lin(int n, int scale_on, const T *__restrict__ diagB, const T *__restrict__ gc,
                                 T *__restrict__ scale_c, T *__restrict__ b_c) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const T s = scale_on ? (T)(1.0 / (DBL_EPSILON + sqrt((double)diagB[i]))) : T(1);
  scale_c[i] = s;
  b_c[i] = s * gc[i];
}

// Deterministic sum of per-tile partials (single CTA).
__global__ void k_sum_partials(const double *__restrict__ part, int n, double *__restrict__ out, int out_idx) {
  __shared__ double shd[32];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += part[i];
  const double tot = block_sum<double>(acc, shd);
  if (threadIdx.x == 0) out[out_idx] = tot;
}

// ---------------------------------------------------------------------------------------------
// damping (hessian.hpp:146-175): d + mu*clamp(d, 1e-6, 1e32)  or  d + mu
// ---------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ T damp_value(T d, T mu, int use_identity) {
  if (use_identity) return (T)((double)d + (double)mu);
  const double dd = (double)d;
  return (T)(dd + mu * fmin(fmax(dd, 1.0e-6), 1.0e32));
}

// K3a: per point — scales, b_p, W = D (D C D + damping)^-1 D, h = W g.
//      replaces execute_block_diagonal_inversion (schur.hpp:1067-1114, cuBLAS matinvBatched 3x3).
template <typename T>
__global__ void k_point_prepare(int Np, int scale_on, T mu, int use_identity, const T *__restrict__ Cg,
                                T *__restrict__ scale_p, T *__restrict__ b_p, T *__restrict__ W, T *__restrict__ h,
                                int write_lin) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= Np) return;
  const T *cg = Cg + (int64_t)p * 9;
  const T c00 = cg[0], c01 = cg[1], c02 = cg[2], c11 = cg[3], c12 = cg[4], c22 = cg[5];
  const T g0 = cg[6], g1 = cg[7], g2 = cg[8];
  T s0 = T(1), s1 = T(1), s2 = T(1);
  if (scale_on) {
    s0 = (T)(1.0 / (DBL_EPSILON + sqrt((double)c00)));
    s1 = (T)(1.0 / (DBL_EPSILON + sqrt((double)c11)));
    s2 = (T)(1.0 / (DBL_EPSILON + sqrt((double)c22)));
  }
  if (write_lin) {
    scale_p[3 * (int64_t)p] = s0; scale_p[3 * (int64_t)p + 1] = s1; scale_p[3 * (int64_t)p + 2] = s2;
    b_p[3 * (int64_t)p] = s0 * g0; b_p[3 * (int64_t)p + 1] = s1 * g1; b_p[3 * (int64_t)p + 2] = s2 * g2;
  }
  // scaled, damped C
  const T a00 = damp_value<T>(s0 * s0 * c00, mu, use_identity);
  const T a11 = damp_value<T>(s1 * s1 * c11, mu, use_identity);
  const T a22 = damp_value<T>(s2 * s2 * c22, mu, use_identity);
  const T a01 = s0 * s1 * c01, a02 = s0 * s2 * c02, a12 = s1 * s2 * c12;
  // symmetric 3x3 inverse by cofactors
  const T m00 = a11 * a22 - a12 * a12, m01 = a02 * a12 - a01 * a22, m02 = a01 * a12 - a02 * a11;
  const T m11 = a00 * a22 - a02 * a02, m12 = a01 * a02 - a00 * a12, m22 = a00 * a11 - a01 * a01;
  const T det = a00 * m00 + a01 * m01 + a02 * m02;
  const T id = T(1) / det;
  const T w00 = s0 * s0 * m00 * id, w01 = s0 * s1 * m01 * id, w02 = s0 * s2 * m02 * id;
  const T w11 = s1 * s1 * m11 * id, w12 = s1 * s2 * m12 * id, w22 = s2 * s2 * m22 * id;
  T *w = W + (int64_t)p * 6;
  w[0] = w00; w[1] = w01; w[2] = w02; w[3] = w11; w[4] = w12; w[5] = w22;
  h[3 * (int64_t)p] = w00 * g0 + w01 * g1 + w02 * g2;
  h[3 * (int64_t)p + 1] = w01 * g0 + w11 * g1 + w12 * g2;
  h[3 * (int64_t)p + 2] = w02 * g0 + w12 * g1 + w22 * g2;
}

template <typename T, typename S>
__device__ __forceinline__ void load_J(const TileStruct &ts, const typename V2<S>::type *__restrict__ Jc,
                                       const typename V2<S>::type *__restrict__ Jp, int64_t o, T *jc, T *jp) {
#pragma unroll
  for (int j = 0; j < 9; j++) {
    const typename V2<S>::type v = Jc[(int64_t)j * ts.Mpad + o];
    jc[2 * j] = (T)v.x;
    jc[2 * j + 1] = (T)v.y;
  }
#pragma unroll
  for (int j = 0; j < 3; j++) {
    const typename V2<S>::type v = Jp[(int64_t)j * ts.Mpad + o];
    jp[2 * j] = (T)v.x;
    jp[2 * j + 1] = (T)v.y;
  }
}

// K3b: per tile — camera-side partials of the Schur diagonal blocks and of the reduced right-hand side:
//   A_c = sum_o Jc^T (I - N_o) Jc   (N_o = Jp W Jp^T; equals B_c - sum E W E^T restricted to the diagonal)
//   u_c = sum_o Jc^T Jp h_p
// replaces execute_schur_multiplication on the diagonal pairs + execute_b_Schur_computation
// (schur.hpp:649-734, 901-920) and the block copy of block_jacobi_schur.hpp:126-137.
template <typename T, typename S>
__global__ void __launch_bounds__(TILE)
k_prepare_tiles(TileStruct ts, const typename V2<S>::type *__restrict__ Jc, const typename V2<S>::type *__restrict__ Jp,
                const T *__restrict__ W, const T *__restrict__ h, T *__restrict__ part /*[nseg][54]*/) {
  __shared__ T sv[TILE * 9];
  const int tile = blockIdx.x, t = threadIdx.x;
  const int o0 = ts.tile_obs[tile], n = ts.tile_obs[tile + 1] - o0;
  const int64_t o = (int64_t)o0 + t;
  const bool active = t < n;
  T jc[18], K[18], q0 = T(0), q1 = T(0);
  int rank = 0;
#pragma unroll
  for (int j = 0; j < 18; j++) { jc[j] = T(0); K[j] = T(0); }
  if (active) {
    T jp[6];
    load_J<T, S>(ts, Jc, Jp, o, jc, jp);
    rank = ts.rank[o];
    const int p = ts.pt_idx[o];
    const T *w = W + (int64_t)p * 6;
    const T w00 = w[0], w01 = w[1], w02 = w[2], w11 = w[3], w12 = w[4], w22 = w[5];
    // rows of Jp: a = (jp[0], jp[2], jp[4]), b = (jp[1], jp[3], jp[5])
    const T wa0 = w00 * jp[0] + w01 * jp[2] + w02 * jp[4];
    const T wa1 = w01 * jp[0] + w11 * jp[2] + w12 * jp[4];
    const T wa2 = w02 * jp[0] + w12 * jp[2] + w22 * jp[4];
    const T wb0 = w00 * jp[1] + w01 * jp[3] + w02 * jp[5];
    const T wb1 = w01 * jp[1] + w11 * jp[3] + w12 * jp[5];
    const T wb2 = w02 * jp[1] + w12 * jp[3] + w22 * jp[5];
    const T n00 = jp[0] * wa0 + jp[2] * wa1 + jp[4] * wa2;
    const T n01 = jp[0] * wb0 + jp[2] * wb1 + jp[4] * wb2;
    const T n11 = jp[1] * wb0 + jp[3] * wb1 + jp[5] * wb2;
    const T m00 = T(1) - n00, m01 = -n01, m11 = T(1) - n11;
#pragma unroll
    for (int j = 0; j < 9; j++) {
      K[2 * j] = m00 * jc[2 * j] + m01 * jc[2 * j + 1];
      K[2 * j + 1] = m01 * jc[2 * j] + m11 * jc[2 * j + 1];
    }
    const T *hp = h + (int64_t)p * 3;
    q0 = jp[0] * hp[0] + jp[2] * hp[1] + jp[4] * hp[2];
    q1 = jp[1] * hp[0] + jp[3] * hp[1] + jp[5] * hp[2];
  }
  const int sg = ts.tile_seg[tile], nsg = ts.tile_seg[tile + 1] - sg;
  // 45 upper entries A(i,j), i <= j, row-wise: (0,0..8), (1,1..8), ... packed index idx; 5 groups of 9
  T v[9];
  int gi = 0, gcount = 0;
#pragma unroll
  for (int i = 0; i < 9; i++) {
#pragma unroll
    for (int j = i; j < 9; j++) {
      v[gcount] = jc[2 * i] * K[2 * j] + jc[2 * i + 1] * K[2 * j + 1];
      gcount++;
      if (gcount == 9) {
        tile_cam_reduce<T>(v, active, rank, o0, sg, nsg, ts.seg_begin, sv, part, 54, gi * 9);
        gcount = 0;
        gi++;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 9; k++) v[k] = jc[2 * k] * q0 + jc[2 * k + 1] * q1;
  tile_cam_reduce<T>(v, active, rank, o0, sg, nsg, ts.seg_begin, sv, part, 54, 45);
}

// In-place Gauss-Jordan inverse of a 9x9 matrix in shared memory by one thread (partial pivoting).
template <typename T> __device__ void invert9(T *A /*[81] col-major*/, T *Ai /*[81]*/) {
  for (int i = 0; i < 81; i++) Ai[i] = T(0);
  for (int i = 0; i < 9; i++) Ai[10 * i] = T(1);
  for (int c = 0; c < 9; c++) {
    int piv = c;
    T best = fabs(A[c + 9 * c]);
    for (int i = c + 1; i < 9; i++) {
      const T v = fabs(A[i + 9 * c]);
      if (v > best) { best = v; piv = i; }
    }
    if (piv != c)
      for (int j = 0; j < 9; j++) {
        T tmp = A[c + 9 * j]; A[c + 9 * j] = A[piv + 9 * j]; A[piv + 9 * j] = tmp;
        tmp = Ai[c + 9 * j]; Ai[c + 9 * j] = Ai[piv + 9 * j]; Ai[piv + 9 * j] = tmp;
      }
    const T ip = T(1) / A[c + 9 * c];
    for (int j = 0; j < 9; j++) { A[c + 9 * j] *= ip; Ai[c + 9 * j] *= ip; }
    for (int i = 0; i < 9; i++) {
      if (i == c) continue;
      const T f = A[i + 9 * c];
      for (int j = 0; j < 9; j++) { A[i + 9 * j] -= f * A[c + 9 * j]; Ai[i + 9 * j] -= f * Ai[c + 9 * j]; }
    }
  }
}

// K3c: per camera — sum partials (or take the allreduced sums), scale, damp, invert.
//   S~_cc = D A D with diagonal + (damp(B~_kk) - B~_kk);  Minv = S~_cc^-1  (block_jacobi_schur.hpp:139-147)
//   b_S   = D (g_c - u_c)                                  (schur.hpp:901-920)
//   dterm = damp(B~_kk) - B~_kk  (added to S p on the diagonal)
template <typename T>
__global__ void __launch_bounds__(288)
k_cam_reduce_prepare(TileStruct ts, const T *__restrict__ part, int from_sums, T *__restrict__ sums /*[Nc][54]*/,
                     int finish, T mu, int use_identity, const T *__restrict__ diagB, const T *__restrict__ gc,
                     const T *__restrict__ scale_c, T *__restrict__ Sdiag /*[Nc][81]*/, T *__restrict__ Minv,
                     T *__restrict__ bS, T *__restrict__ dterm) {
  __shared__ T sh[32 * 9];
  __shared__ T out[54];
  __shared__ T A[81], Ai[81];
  const int c = blockIdx.x, t = threadIdx.x;
  if (!from_sums) {
    cam_gather<T>(ts, c, part, 54, 6, sh, out);
    if (t < 54) sums[(int64_t)c * 54 + t] = out[t];
  } else {
    if (t < 54) out[t] = sums[(int64_t)c * 54 + t];
    __syncthreads();
  }
  if (!finish) return;
  if (t < 81) {
    const int i = t % 9, j = t / 9;
    const int a = i < j ? i : j, b = i < j ? j : i;
    const int idx = a * 9 - (a * (a - 1)) / 2 + (b - a); // packed upper index of (a,b), row-wise
    const T si = scale_c[c * 9 + i], sj = scale_c[c * 9 + j];
    T val = si * sj * out[idx];
    if (i == j) {
      const T bt = si * si * diagB[c * 9 + i];
      const T dt = damp_value<T>(bt, mu, use_identity) - bt;
      val += dt;
      dterm[c * 9 + i] = dt;
      bS[c * 9 + i] = si * (gc[c * 9 + i] - out[45 + i]);
    }
    A[i + 9 * j] = val;
    Sdiag[(int64_t)c * 81 + i + 9 * j] = val;
  }
  __syncthreads();
  if (t == 0) invert9<T>(A, Ai);
  __syncthreads();
  if (t < 81) Minv[(int64_t)c * 81 + t] = Ai[t];
}

// ---------------------------------------------------------------------------------------------
// K4: matrix-free Schur product, per tile.  xs = D_c x (10-padded rows).
//   y_o = Jc x_c ; t_p = sum_o Jp^T y_o ; w_p = W_p t_p ; z_o = Jp w_p ; v_o = Jc^T (y_o - z_o)
//   part[segment] = sum over the segment of v_o        ( = (B - E W E^T) x restricted to the tile )
// replaces execute_schur_vector_multiply (schur.hpp:347-393) on an explicit S.
// mode 1 (back-substitution, schur.hpp:279-302 + ops/update.hpp:9-31): instead of z/v, finish per point
//   x~_p = (h_p - W_p t_p) / s_p ; delta_p = x~_p s_p ; backup and update the point ; rho partial.
// ---------------------------------------------------------------------------------------------
template <typename T, typename S, int MODE>
__global__ void __launch_bounds__(TILE)
k_schur_tiles(TileStruct ts, const typename V2<S>::type *__restrict__ Jc, const typename V2<S>::type *__restrict__ Jp,
              const T *__restrict__ W, const T *__restrict__ xs, T *__restrict__ part /*[nseg][9]*/,
              // MODE 1 only:
              const T *__restrict__ h, const T *__restrict__ scale_p, const T *__restrict__ b_p, T mu,
              T *__restrict__ pts, T *__restrict__ pts_bak, T *__restrict__ delta_p, double *__restrict__ rho_part,
              int apply, const int *__restrict__ done_flag) {
  __shared__ T sv[TILE * 9];
  __shared__ T sw[TILE * 3];
  __shared__ double shd[32];
  if (done_flag && *done_flag) return; // PCG already stopped: nothing to do (uniform across the grid)
  const int tile = blockIdx.x, t = threadIdx.x;
  const int o0 = ts.tile_obs[tile], n = ts.tile_obs[tile + 1] - o0;
  const int64_t o = (int64_t)o0 + t;
  const bool active = t < n;
  const int p0 = ts.tile_pt[tile], npt = ts.tile_pt[tile + 1] - p0;
  T jc[18], jp[6], y0 = T(0), y1 = T(0);
  int rank = 0, plocal = 0;
  if (active) {
    load_J<T, S>(ts, Jc, Jp, o, jc, jp);
    const int c = ts.cam_idx[o];
    plocal = ts.pt_idx[o] - p0;
    rank = ts.rank[o];
    T x[10];
    load_cam<T>(xs, c, x);
#pragma unroll
    for (int j = 0; j < 9; j++) {
      y0 += jc[2 * j] * x[j];
      y1 += jc[2 * j + 1] * x[j];
    }
    sv[t * 3 + 0] = jp[0] * y0 + jp[1] * y1;
    sv[t * 3 + 1] = jp[2] * y0 + jp[3] * y1;
    sv[t * 3 + 2] = jp[4] * y0 + jp[5] * y1;
  }
  __syncthreads();
  double rho = 0.0;
  if (t < npt) {
    const int p = p0 + t;
    const int b = ts.pptr[p] - o0, e = ts.pptr[p + 1] - o0;
    T t0 = T(0), t1 = T(0), t2 = T(0);
    for (int row = b; row < e; row++) {
      t0 += sv[row * 3];
      t1 += sv[row * 3 + 1];
      t2 += sv[row * 3 + 2];
    }
    const T *w = W + (int64_t)p * 6;
    const T w0 = w[0] * t0 + w[1] * t1 + w[2] * t2;
    const T w1 = w[1] * t0 + w[3] * t1 + w[4] * t2;
    const T w2 = w[2] * t0 + w[4] * t1 + w[5] * t2;
    if (MODE == 0) {
      sw[t * 3] = w0; sw[t * 3 + 1] = w1; sw[t * 3 + 2] = w2;
    } else {
      const T wv[3] = {w0, w1, w2};
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const int64_t i = 3 * (int64_t)p + k;
        const T s = scale_p[i];
        const T xt = (h[i] - wv[k]) / s; // scaled-space step of the point
        delta_p[i] = xt;
        rho += (double)(xt * (mu * xt + b_p[i]));
        if (apply) {
          const T old = pts[i];
          pts_bak[i] = old;
          pts[i] = old + xt * s;
        }
      }
    }
  }
  if (MODE == 1) {
    const double tot = block_sum<double>(rho, shd);
    if (t == 0) rho_part[tile] = tot;
    return;
  }
  __syncthreads();
  T v[9];
  if (active) {
    const T z0 = jp[0] * sw[plocal * 3] + jp[2] * sw[plocal * 3 + 1] + jp[4] * sw[plocal * 3 + 2];
    const T z1 = jp[1] * sw[plocal * 3] + jp[3] * sw[plocal * 3 + 1] + jp[5] * sw[plocal * 3 + 2];
    const T d0 = y0 - z0, d1 = y1 - z1;
#pragma unroll
    for (int k = 0; k < 9; k++) v[k] = jc[2 * k] * d0 + jc[2 * k + 1] * d1;
  }
  __syncthreads(); // sv is reused by the camera reduction
  const int sg = ts.tile_seg[tile], nsg = ts.tile_seg[tile + 1] - sg;
  tile_cam_reduce<T>(v, active, rank, o0, sg, nsg, ts.seg_begin, sv, part, 9, 0);
}

// Camera side of the product: Ap_raw = D_c * sum(partials)   (the damping term is added by the PCG update).
template <typename T>
__global__ void __launch_bounds__(288)
k_cam_reduce_spmv(TileStruct ts, const T *__restrict__ part, const T *__restrict__ scale_c, T *__restrict__ Ap,
                  const int *__restrict__ done_flag) {
  __shared__ T sh[32 * 9];
  __shared__ T out[9];
  if (done_flag && *done_flag) return;
  const int c = blockIdx.x;
  cam_gather<T>(ts, c, part, 9, 1, sh, out);
  if (threadIdx.x < 9) Ap[c * 9 + threadIdx.x] = scale_c[c * 9 + threadIdx.x] * out[threadIdx.x];
}

// ---------------------------------------------------------------------------------------------
// PCG on the reduced camera system (solver/pcg_schur.hpp:79-168).  Scalars stay on the device;
// every CTA recomputes the two dot products in the same fixed order, so all CTAs (and all ranks)
// take identical decisions without a broadcast.  State is ping-ponged: kernel k reads st[k], CTA 0
// writes st[k+1].
// ---------------------------------------------------------------------------------------------
template <typename T> struct PcgState {
  T rz, rz0, alpha, beta, denom;
  int iter, done, reason, pad;
};

constexpr int PCG_CAMS = 32; // cameras per CTA (288 threads)

template <typename T> __device__ __forceinline__ T dot_all(const T *a, const T *b, const T *dterm, int n, T *sh) {
  // sum_i a_i * (b_i + dterm_i a_i); dterm may be null
  T acc = T(0);
  if (dterm)
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc += a[i] * (b[i] + dterm[i] * a[i]);
  else
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc += a[i] * b[i];
  return block_sum<T>(acc, sh);
}

// x = 0 ; r = bS ; z = Minv r ; p = z ; xs = D p.
template <typename T>
__global__ void __launch_bounds__(288)
k_pcg_init(int Nc, const T *__restrict__ bS, const T *__restrict__ Minv, const T *__restrict__ scale_c,
           T *__restrict__ x, T *__restrict__ r, T *__restrict__ z, T *__restrict__ p, T *__restrict__ xs) {
  __shared__ T sr[288];
  const int t = threadIdx.x, c = blockIdx.x * PCG_CAMS + t / 9, k = t % 9;
  const bool ok = c < Nc;
  const int i = c * 9 + k;
  sr[t] = ok ? bS[i] : T(0);
  __syncthreads();
  if (!ok) return;
  T acc = T(0);
  const T *m = Minv + (int64_t)c * 81;
  const T *rc = sr + (t / 9) * 9;
#pragma unroll
  for (int j = 0; j < 9; j++) acc += m[k + 9 * j] * rc[j];
  x[i] = T(0);
  r[i] = sr[t];
  z[i] = acc;
  p[i] = acc;
  xs[c * CAM_STRIDE + k] = scale_c[i] * acc;
  if (k == 0) xs[c * CAM_STRIDE + 9] = T(0);
}
template <typename T>
__global__ void __launch_bounds__(1024)
k_pcg_init_state(int n, const T *__restrict__ r, const T *__restrict__ z, PcgState<T> *st) {
  __shared__ T sh[32];
  const T rz = dot_all<T>(r, z, nullptr, n, sh);
  if (threadIdx.x == 0) {
    PcgState<T> s;
    s.rz = rz; s.rz0 = (T)INFINITY; s.alpha = T(0); s.beta = T(0); s.denom = T(0);
    s.iter = 0; s.done = 0; s.reason = 0; s.pad = 0;
    st[0] = s;
  }
}

// first half of an iteration (after Ap_raw = S_undamped p): denom, alpha, x/r update, z = Minv r
template <typename T>
__global__ void __launch_bounds__(288)
k_pcg_update1(int Nc, const PcgState<T> *__restrict__ sin, PcgState<T> *__restrict__ sout,
              const T *__restrict__ Ap_raw, const T *__restrict__ dterm, const T *__restrict__ Minv,
              T *__restrict__ Ap, T *__restrict__ x, T *__restrict__ xbak, T *__restrict__ r, T *__restrict__ z,
              const T *__restrict__ p, int *done_flag) {
  __shared__ T sh[32];
  __shared__ T sr[288];
  PcgState<T> s = *sin;
  if (s.done) {
    if (blockIdx.x == 0 && threadIdx.x == 0) *sout = s;
    return;
  }
  const int n = Nc * 9;
  if (s.rz == T(0)) { // pcg_schur.hpp:109-111
    if (blockIdx.x == 0 && threadIdx.x == 0) { s.done = 1; s.reason = 3; *sout = s; *done_flag = 1; }
    return;
  }
  const T denom = dot_all<T>(p, Ap_raw, dterm, n, sh);
  if (denom == T(0) || isnan(denom)) { // :120-122
    if (blockIdx.x == 0 && threadIdx.x == 0) { s.done = 1; s.reason = 4; s.denom = denom; *sout = s; *done_flag = 1; }
    return;
  }
  const T alpha = s.rz / denom;
  const int t = threadIdx.x, c = blockIdx.x * PCG_CAMS + t / 9, k = t % 9;
  const bool ok = c < Nc;
  const int i = c * 9 + k;
  T rn = T(0);
  if (ok) {
    const T pi = p[i];
    const T ap = Ap_raw[i] + dterm[i] * pi;
    Ap[i] = ap;
    const T xo = x[i];
    xbak[i] = xo;
    x[i] = alpha * pi + xo;   // ops::axpy_async(x, alpha, p, x)
    rn = -alpha * ap + r[i];  // ops::axpy_async(r, -alpha, Ap, r)
    r[i] = rn;
  }
  sr[t] = rn;
  __syncthreads();
  if (ok) {
    T acc = T(0);
    const T *m = Minv + (int64_t)c * 81;
    const T *rc = sr + (t / 9) * 9;
#pragma unroll
    for (int j = 0; j < 9; j++) acc += m[k + 9 * j] * rc[j];
    z[i] = acc;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    s.alpha = alpha;
    s.denom = denom;
    *sout = s;
  }
}

// second half: rz_new, rejection / convergence tests, beta, p update, xs = D p
template <typename T>
__global__ void __launch_bounds__(288)
k_pcg_update2(int Nc, const PcgState<T> *__restrict__ sin, PcgState<T> *__restrict__ sout, T tol, T ratio,
              int max_iter, const T *__restrict__ scale_c, T *__restrict__ x, const T *__restrict__ xbak,
              const T *__restrict__ r, const T *__restrict__ z, T *__restrict__ p, T *__restrict__ xs,
              int *done_flag) {
  __shared__ T sh[32];
  PcgState<T> s = *sin;
  if (s.done) {
    if (blockIdx.x == 0 && threadIdx.x == 0) *sout = s;
    return;
  }
  const int n = Nc * 9;
  const T rzn = dot_all<T>(r, z, nullptr, n, sh);
  const int t = threadIdx.x, c = blockIdx.x * PCG_CAMS + t / 9, k = t % 9;
  const bool ok = c < Nc;
  const int i = c * 9 + k;
  const bool leader = blockIdx.x == 0 && threadIdx.x == 0;
  s.iter += 1;
  if (fabs(rzn) > ratio * s.rz0 || isnan(rzn)) { // :144-148
    if (ok) x[i] = xbak[i];
    if (leader) { s.done = 1; s.reason = 2; s.rz = rzn; *sout = s; *done_flag = 1; }
    return;
  }
  s.rz0 = fmin(s.rz0, fabs(rzn));
  const T beta = rzn / s.rz;
  s.beta = beta;
  s.rz = rzn;
  if (ok) {
    const T pn = beta * p[i] + z[i]; // ops::axpy_async(p, beta, p, z)
    p[i] = pn;
    xs[c * CAM_STRIDE + k] = scale_c[i] * pn;
  }
  if (fabs(rzn) < tol) { s.done = 1; s.reason = 1; }
  else if (s.iter >= max_iter) { s.done = 1; s.reason = 0; }
  if (leader) {
    *sout = s;
    if (s.done) *done_flag = 1;
  }
}

// xs = D_c x (for the back-substitution) ; also camera update + rho partial (ops/update.hpp:9-31,
// levenberg_marquardt.hpp:34-41).  Single pass over the 9 Nc camera scalars.
template <typename T>
__global__ void k_cam_step(int n, const T *__restrict__ x, const T *__restrict__ scale_c, const T *__restrict__ b_c,
                           T mu, T *__restrict__ xs, T *__restrict__ cams, T *__restrict__ cams_bak,
                           T *__restrict__ delta_c, double *__restrict__ rho_part, int apply) {
  __shared__ double shd[32];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  double rho = 0.0;
  if (i < n) {
    const int c = i / 9, k = i - 9 * c;
    const T xt = x[i], s = scale_c[i];
    const T d = xt * s;
    xs[c * CAM_STRIDE + k] = d;
    delta_c[i] = xt;
    rho = (double)(xt * (mu * xt + b_c[i]));
    if (apply) {
      const T old = cams[c * CAM_STRIDE + k];
      cams_bak[c * CAM_STRIDE + k] = old;
      cams[c * CAM_STRIDE + k] = old + d;
    }
  }
  const double tot = block_sum<double>(rho, shd);
  if (threadIdx.x == 0) rho_part[blockIdx.x] = tot;
}

// K5 cost: residual only (graph.hpp:221-234 compute_error + chi2)
template <typename T>
__global__ void __launch_bounds__(TILE)
k_cost_tiles(TileStruct ts, const T *__restrict__ cams, const T *__restrict__ pts,
             const typename V2<T>::type *__restrict__ obs, double *__restrict__ cost_part) {
  __shared__ double shd[32];
  const int tile = blockIdx.x, t = threadIdx.x;
  const int o0 = ts.tile_obs[tile], n = ts.tile_obs[tile + 1] - o0;
  const int64_t o = (int64_t)o0 + t;
  double cost = 0.0;
  if (t < n) {
    const int c = ts.cam_idx[o], p = ts.pt_idx[o];
    T cam[10], X[3], ob[2], r[2];
    load_cam<T>(cams, c, cam);
    X[0] = pts[3 * (int64_t)p];
    X[1] = pts[3 * (int64_t)p + 1];
    X[2] = pts[3 * (int64_t)p + 2];
    const typename V2<T>::type ov = obs[o];
    ob[0] = ov.x;
    ob[1] = ov.y;
    bal_residual<T>(cam, X, ob, r);
    cost = (double)(r[0] * r[0] + r[1] * r[1]);
  }
  const double tot = block_sum<double>(cost, shd);
  if (t == 0) cost_part[tile] = tot;
}

// ---------------------------------------------------------------------------------------------
// parity exports
// ---------------------------------------------------------------------------------------------
// Scaled, undamped Hessian values in the reference layout (hessian.hpp:257-288):
// [B_0 .. B_{Nc-1}] then per point [E_{c1,p} E_{c2,p} ... C_p], blocks column-major.
// E and C per observation / point here; B via k_hessian_B.
template <typename T, typename S>
__global__ void k_hessian_EC(TileStruct ts, const typename V2<S>::type *__restrict__ Jc,
                             const typename V2<S>::type *__restrict__ Jp, const T *__restrict__ Cg,
                             const T *__restrict__ scale_c, const T *__restrict__ scale_p, S *__restrict__ vals) {
  const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (o < ts.M) {
    T jc[18], jp[6];
    load_J<T, S>(ts, Jc, Jp, o, jc, jp);
    const int c = ts.cam_idx[o], p = ts.pt_idx[o];
    // offset: 81 Nc + 27 o + 9 p  (every earlier point contributes its E blocks and one C block)
    S *dst = vals + 81 * (int64_t)ts.Nc + 27 * o + 9 * (int64_t)p;
    for (int j = 0; j < 3; j++)
      for (int i = 0; i < 9; i++) {
        const T a0 = (T)(S)(jc[2 * i] * scale_c[c * 9 + i]), a1 = (T)(S)(jc[2 * i + 1] * scale_c[c * 9 + i]);
        const T b0 = (T)(S)(jp[2 * j] * scale_p[3 * (int64_t)p + j]), b1 = (T)(S)(jp[2 * j + 1] * scale_p[3 * (int64_t)p + j]);
        dst[i + 9 * j] = (S)(a0 * b0 + a1 * b1);
      }
  }
  if (o < ts.Np) {
    const int64_t p = o;
    const T *cg = Cg + p * 9;
    const T s[3] = {scale_p[3 * p], scale_p[3 * p + 1], scale_p[3 * p + 2]};
    const T cf[9] = {cg[0], cg[1], cg[2], cg[1], cg[3], cg[4], cg[2], cg[4], cg[5]};
    S *dst = vals + 81 * (int64_t)ts.Nc + 27 * (int64_t)ts.pptr[p + 1] + 9 * p;
    for (int k = 0; k < 9; k++) dst[k] = (S)(s[k % 3] * s[k / 3] * cf[k]);
  }
}
// B blocks: one CTA per camera, gathers through the camera -> segment lists (export only; not a hot path)
template <typename T, typename S>
__global__ void k_hessian_B(TileStruct ts, const typename V2<S>::type *__restrict__ Jc, const T *__restrict__ scale_c,
                            S *__restrict__ vals) {
  const int c = blockIdx.x, t = threadIdx.x; // 81 threads
  if (t >= 81) return;
  const int i = t % 9, j = t / 9;
  T acc = T(0);
  for (int q = ts.cam_seg_ptr[c]; q < ts.cam_seg_ptr[c + 1]; q++) {
    const int s = ts.cam_seg_list[q];
    // observations of the segment: those whose rank falls in [seg_begin[s], seg_begin[s+1]) — the export walks
    // the tile to find them (slow, test-only)
    int lo = 0, hi = ts.ntiles; // tile containing the segment: largest tile with tile_seg <= s
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (ts.tile_seg[mid] <= s) lo = mid; else hi = mid;
    }
    const int o0 = ts.tile_obs[lo], n = ts.tile_obs[lo + 1] - o0;
    const int rb = ts.seg_begin[s] - o0, re = ts.seg_begin[s + 1] - o0;
    for (int u = 0; u < n; u++) {
      const int rk = ts.rank[o0 + u];
      if (rk >= rb && rk < re) {
        const int64_t o = o0 + u;
        const typename V2<S>::type a = Jc[(int64_t)i * ts.Mpad + o], b = Jc[(int64_t)j * ts.Mpad + o];
        const T a0 = (T)a.x, a1 = (T)a.y, b0 = (T)b.x, b1 = (T)b.y;
        acc += a0 * b0 + a1 * b1;
      }
    }
  }
  vals[(int64_t)c * 81 + i + 9 * j] = (S)(scale_c[c * 9 + i] * scale_c[c * 9 + j] * acc);
}

} // namespace gb

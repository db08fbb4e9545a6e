// kernels.cuh — the sm_100a kernels of the LM inner loop.
//
// Layout (HBM).  Observations are sorted by (point, camera) and cut into TILES of whole points; a tile owns
// 256 storage slots (slot = tile * 256 + position, unused slots are padding with zero Jacobians), one thread
// per slot.  Consecutive tiles form a SUPER-TILE owned by one CTA (structure.hpp).
//   J     tile-major: J[tile][12 planes][256] of vec2<S>; planes 0-8 = columns of the 2x9 camera Jacobian,
//         planes 9-11 = columns of the 2x3 point Jacobian.  Every plane access is a coalesced 16-byte stream.
//   r,obs vec2<T> per slot.
//   per point:  Cg[9] = {C00,C01,C02,C11,C12,C22, g0,g1,g2}  (C = sum Jp^T Jp, g = -sum Jp^T r, unscaled),
//               W[6] (= D_p (D_p C D_p + damping)^-1 D_p), h[3] = W g.
//   per camera: 10-padded rows (80 B in FP64) so that a gather is five 16-byte loads.
//
// Reductions are atomic-free and deterministic:
//   by point  - a point's observations are contiguous inside one tile: staged in shared memory, summed
//               sequentially by one thread per (point, component);
//   by camera - the slots of a tile are in (camera, observation) order, so a camera SEGMENT is a run of adjacent
//               threads.  Threads stage their 9-vectors in shared memory at their own row; one
//               thread per (segment, component) sums the segment and adds it to the super-tile's accumulator
//               row of that camera in shared memory (exactly one writer per row and tile, tiles in order).
//               At the end the CTA writes its rows; a per-camera kernel sums the rows of all super-tiles in
//               ascending order (camera -> row CSR).
// All arithmetic on the Jacobians is done in the UNSCALED space; the Jacobi scaling of the reference
// (graph.hpp:254-281) is applied as D_c / D_p on the camera- and point-sized quantities, which is the
// same algebra: J~ = J D  =>  J~^T J~ = D J^T J D.
//
// Kernels of one LM iteration, in launch order (DESIGN.md section 3 has bytes, bounds and measured times):
//   k_cam_precompute, k_linearize<EXT>   factor evaluation (built-in BAL model or the caller's kernel), whitening by
//                                        loss / precision, J + residual store, C and g per point, diag(B) and g_c partials
//   k_cam_reduce_lin, k_point_prepare    Jacobi scales, b; W = D (D C D + damping)^-1 D, h = W g
//   k_prepare_cams, k_cam_reduce_prepare diagonal blocks of S and b_S by a camera-major gather; 9x9 inverses
//   k_pcg_solve (pcg_solve.cuh)          the whole PCG solve in ONE persistent cooperative launch: per iteration the
//                                        matrix-free (B - E W E^T) p on the TMA pipeline (product_tile below), row sums,
//                                        multi-GPU exchange (p2p.cuh), dots and updates, two grid barriers
//   k_cam_step, k_backsubst_points       step of the cameras; back-substitution and step of the points from the point sums the
//                                        solve kernel kept (no pass over J); k_backsubst_tiles streams J instead (long solves,
//                                        explicit / direct Schur); rho partials
//   k_cost_tiles, k_sum_partials3        cost at the trial point, cost + rho sums; k_store_host hands them to the host
//   k_copy                               restore after a rejected step
// Other entry points: k_schur_product2 (one product per launch: exports, <FULL> for the full-system PCG solver, NCCL
// fallback) + k_cam_reduce_spmv, k_pcg_init*, k_pcg_update (NCCL fallback), k_full_* (full-system PCG solver: a
// device-resident loop, scalars and stop rules in a FullState), k_frag_sum / k_frag_dots (second level of the per-point sums
// of long tracks), k_hessian_export, k_scatter_slots, k_p2p_push / k_p2p_sum (generic exchange).
#pragma once
#include <cfloat>
#include <cstdint>
#include <cooperative_groups.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "bal_math.cuh"
#include "p2p.cuh"
#include "structure.hpp"

namespace gb {

constexpr int CAM_STRIDE = 10; // padded camera row
constexpr int NPLANES = 12;
// dynamic shared memory of the super-tile kernels, in elements of T
#ifndef LIN_MIN_BLOCKS
#define LIN_MIN_BLOCKS 2
#endif
constexpr int COST_TILES = 4; // tiles per CTA of k_cost_tiles
constexpr int LIN_STR = 19; // staging row stride of k_linearize (18 values; odd stride: conflict-free 64-bit rows)
constexpr int LIN_PTS = 128 * 3; // staged point coordinates of a tile (<= 128 points)
constexpr int LIN_CXS = CAMX + 1; // row stride of the staged camera table: odd, so that the rows of different cameras hit different banks
template <typename T> constexpr int smem_lin_bytes() { // staging, accumulator rows, tile record, camera table, tile points
  return (TILE * LIN_STR + SLOT_CAP * 18) * (int)sizeof(T) + 16 + REC_BYTES + (SLOT_CAP * LIN_CXS + LIN_PTS) * (int)sizeof(T);
}
// per-point W row stride: 6 values, padded to 8 in FP32 so that a tile's rows start 16-byte aligned (TMA)
template <typename T> struct WST { static constexpr int value = sizeof(T) == 4 ? 8 : 6; };
constexpr int HST = 4; // per-point h row stride (3 values + pad): 16-byte aligned rows in FP32 and FP64

template <typename T> struct V2;
template <> struct V2<double> {
  using type = double2;
  static __device__ __forceinline__ type make(double a, double b) { return make_double2(a, b); }
};
template <> struct V2<float> {
  using type = float2;
  static __device__ __forceinline__ type make(float a, float b) { return make_float2(a, b); }
};
// bf16 Jacobian storage (the reference's low-precision S, types.hpp:10-19; examples/bal.cu:186-236): 4 bytes per pair
template <> struct V2<__nv_bfloat16> {
  using type = __nv_bfloat162;
  static __device__ __forceinline__ type make(__nv_bfloat16 a, __nv_bfloat16 b) { return __halves2bfloat162(a, b); }
};
template <typename S> struct IsLowPrecision { static constexpr bool value = false; };
template <> struct IsLowPrecision<__nv_bfloat16> { static constexpr bool value = true; };

struct DevStruct {
  int64_t M, Mstore;
  int32_t Nc, Np, ntiles, nst, nrows, pad;
  const TileMeta *tmeta;     // [ntiles]
  const uint32_t *ometa;     // [Mstore]  cslot:16 | position in point order:8 | point-in-tile:8
  const unsigned char *trec; // [ntiles][REC_BYTES] packed per-tile record (ometa | seg_tab | pt_tab | meta)
  const int32_t *tile_cam;   // [Mstore] camera of each storage slot (one-CTA-per-tile kernels)
  const int32_t *st_tile;    // [nst+1]
  const int32_t *st_row;     // [nst+1] first partial row of the super-tile
  const int32_t *row_cam;    // [nrows] camera of each partial row
  const int32_t *cam_row_ptr;  // [Nc+1] camera -> its contiguous rows in the partial buffers (ascending super-tile)
  const int32_t *row_out;      // [nrows] super-tile row (st_row + slot) -> position in the partial buffers
  const unsigned char *strec;  // [nst][STREC_BYTES] packed per-super-tile record (k_pcg_solve)
  const int32_t *cta_st;       // [ncta+1] super-tile ranges of the persistent CTAs
  int32_t ncta, pad2;
  const int32_t *cm_slot, *cm_pt;  // [M] camera-major observation -> storage slot / point
  const int32_t *ch_ptr;           // [nchunks+1] chunk -> camera-major observation range
  const int32_t *cam_ch_ptr;       // [Nc+1] camera -> its chunk rows
  int32_t nchunks, pad3;
  const int32_t *slot_of_obs; // [M] sorted observation -> storage slot (exports)
  const int32_t *cam_idx, *pt_idx, *pptr;    // sorted observations (exports)
  // long tracks (structure.hpp): a point with more observations than a tile holds is cut into fragment tiles (np = 1,
  // TileMeta::frag != 0); its per-point sums have two levels
  int32_t nfrag, nheavy;
  const int32_t *hv_pt, *hv_ptr; // [nheavy] the points, [nheavy+1] their fragment ranges
  void *frag_part;               // [nfrag][9] of T: per-fragment sums (C | g in k_linearize; point part of the full-system product)
  void *frag_t;                  // [nfrag][3] of T: sum_o Jp^T Jc x_c over ALL observations of the fragment's point (k_frag_dots)
  // per-iteration point sums of k_pcg_solve kept for the back-substitution (k_backsubst_points): tk[k][Np][3] of T holds
  // t_p(p_k) = sum_o Jp^T Jc (D p_k)_c of PCG iteration k, tk_alpha[k] its step length (0: not applied); tk_cap = 0: off
  void *tk, *tk_alpha;
  int32_t tk_cap, pad4;
};

// ---------------------------------------------------------------------------------------------
// TMA (bulk async copy) + mbarrier helpers — sm_90+/sm_100a.  One thread issues, all threads wait on parity.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
  } while (!ok);
}
// global -> shared bulk copy, completion counted in bytes on the mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// the same with an L2 eviction-priority hint: the Jacobian stream (1 GB per pass, read once) is marked evict-first so
// that the small vectors every pass re-reads (W, partial rows, camera vectors: ~70 MB) stay resident in the 126 MB L2
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void bulk_g2s_hint(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
               : "memory");
}
// order generic-proxy accesses to shared memory before later async-proxy (TMA) writes to the same bytes
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// bar.sync among the 256 threads of one worker (ids 1 and 2; id 0 is __syncthreads)
__device__ __forceinline__ void worker_sync(int worker) { asm volatile("bar.sync %0, 256;" ::"r"(worker + 1) : "memory"); }

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

// Stage v[9] at row `rank`, sum every camera segment of the tile and add it to the accumulator row of the
// segment's camera slot.  acc row stride astride, component offset aoff.
template <typename T>
__device__ __forceinline__ void tile_cam_accumulate(const T v[9], int rank, int nseg, const uint32_t *st /*global or shared*/,
                                                    T *sv /*[TILE*9]*/, T *acc, int astride, int aoff) {
#pragma unroll
  for (int k = 0; k < 9; k++) sv[rank * 9 + k] = v[k];
  __syncthreads();
  for (int item = threadIdx.x; item < nseg * 9; item += TILE) {
    const int s = item / 9, k = item - 9 * s;
    const uint32_t e0 = st[s], e1 = st[s + 1];
    const int b = (int)(e0 >> 16), e = (int)(e1 >> 16), cslot = (int)(e0 & 0xffffu);
    T a = T(0);
    for (int row = b; row < e; row++) a += sv[row * 9 + k];
    acc[cslot * astride + aoff + k] += a;
  }
  __syncthreads();
}

// Sum the partial rows of one camera (block per camera, 288 threads = 32 sub-lists x 9 components).
// out[g*9+k] for g < ngroups; deterministic: sub-list i takes rows i, i+32, ...; then 0..31 in order.
template <typename T>
__device__ __forceinline__ void cam_gather(const DevStruct &ds, int c, const T *__restrict__ part, int pstride,
                                           int ngroups, T *sh /*[32*9]*/, T *out /*smem [ngroups*9]*/,
                                           const int32_t *__restrict__ row_ptr = nullptr) {
  const int sub = threadIdx.x / 9, k = threadIdx.x - 9 * sub;
  if (!row_ptr) row_ptr = ds.cam_row_ptr;
  const int b = row_ptr[c], e = row_ptr[c + 1];
  for (int g = 0; g < ngroups; g++) {
    T acc = T(0);
    if (sub < 32)
      for (int i = b + sub; i < e; i += 32) acc += part[(int64_t)i * pstride + g * 9 + k];
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

template <typename T, typename S>
__device__ __forceinline__ void load_J(const typename V2<S>::type *__restrict__ J, int tile, int t, T *jc, T *jp) {
  const typename V2<S>::type *base = J + ((int64_t)tile * NPLANES) * TILE + t;
#pragma unroll
  for (int j = 0; j < 9; j++) {
    const typename V2<S>::type v = base[j * TILE];
    jc[2 * j] = (T)v.x;
    jc[2 * j + 1] = (T)v.y;
  }
#pragma unroll
  for (int j = 0; j < 3; j++) {
    const typename V2<S>::type v = base[(9 + j) * TILE];
    jp[2 * j] = (T)v.x;
    jp[2 * j + 1] = (T)v.y;
  }
}

// per-camera part of the camera model (rotation matrix, sin/cos terms): one thread per camera, run whenever the
// camera parameters change
template <typename T> __global__ void k_cam_precompute(int Nc, const T *__restrict__ cams, T *__restrict__ camx) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= Nc) return;
  T cam[10], cx[CAMX];
  load_cam<T>(cams, c, cam);
  bal_cam_precompute<T>(cam, cx);
#pragma unroll
  for (int i = 0; i < CAMX; i++) camx[(int64_t)c * CAMX + i] = cx[i];
}
template <typename T> __device__ __forceinline__ void load_camx(const T *__restrict__ camx, int c, T *cx);
template <> __device__ __forceinline__ void load_camx<double>(const double *__restrict__ camx, int c, double *cx) {
  const double2 *p = reinterpret_cast<const double2 *>(camx + (int64_t)c * CAMX);
#pragma unroll
  for (int i = 0; i < CAMX / 2; i++) {
    const double2 v = __ldg(p + i);
    cx[2 * i] = v.x;
    cx[2 * i + 1] = v.y;
  }
}
template <> __device__ __forceinline__ void load_camx<float>(const float *__restrict__ camx, int c, float *cx) {
  const float4 *p = reinterpret_cast<const float4 *>(camx + (int64_t)c * CAMX);
#pragma unroll
  for (int i = 0; i < CAMX / 4; i++) {
    const float4 v = __ldg(p + i);
    cx[4 * i] = v.x; cx[4 * i + 1] = v.y; cx[4 * i + 2] = v.z; cx[4 * i + 3] = v.w;
  }
}

// ---------------------------------------------------------------------------------------------
// Loss and precision matrix of a factor (ops/chi2.hpp:9-44, loss.hpp:15-51, factor.hpp:397-405).
//   chi2_f = loss(r^T P r), dL = loss'(r^T P r);  H += dL J^T P J,  b -= dL J^T P r  (ops/hessian.hpp:58-76,
//   ops/linearize.hpp:277-302).
// With P = U^T U (U upper triangular, factored on the host when the matrices are set) and w = sqrt(dL) the factor is
// WHITENED once, where it is evaluated:  J' = w U J,  r' = w U r,  so that J'^T J' = dL J^T P J and J'^T r' = dL J^T P r
// and every later kernel (assembly, Schur products, back-substitution) works on J' unchanged.
// Pu = nullptr and loss_kind = 0 (DefaultLoss, P = I) take the original code path.
// ---------------------------------------------------------------------------------------------
struct Robust {
  const void *Pu;  // [Mstore][3] of T: u00, u01, u11 per storage slot, or nullptr
  int loss_kind;   // 0 DefaultLoss, 1 HuberLoss
  double delta;
};
// User-defined factor (FactorTraits::error / ::jacobian, docs/markdown/main.md:284-289; dispatch ops/error.hpp:33-96,
// ops/linearize.hpp:8-138): residuals and column-major 2x9 / 2x3 Jacobians evaluated by the CALLER's kernel into
// caller-order buffers of T; the tile kernels read them through slot_src (storage slot -> caller's factor index).
struct ExtFactor {
  const void *r, *Jc, *Jp;    // [n_obs][2], [n_obs][18], [n_obs][6] of T
  const int32_t *slot_src;    // [Mstore], -1 in padding slots
  const void *rescale;        // k_linearize<RESCALE>: Jacobi scales [9 Nc + 3 Np] of T the stored Jacobians are multiplied by
};
// returns chi2_f; whitens r (2), Jc (18), Jp (6) in place
template <typename T, typename S>
__device__ __forceinline__ T whiten_factor(const Robust &rb, int64_t slot, T *r, T *Jc, T *Jp) {
  T u00 = T(1), u01 = T(0), u11 = T(1);
  if (rb.Pu) {
    const T *u = reinterpret_cast<const T *>(rb.Pu) + 3 * slot;
    u00 = u[0]; u01 = u[1]; u11 = u[2];
  }
  const T e0 = u00 * r[0] + u01 * r[1], e1 = u11 * r[1];
  const T c = e0 * e0 + e1 * e1; // r^T P r
  T chi = c, w = T(1);
  const T delta = (T)rb.delta;
  if (rb.loss_kind == 1 && c > delta * delta) {
    const T sc = sqrt(c);
    chi = T(2) * sc * delta - delta * delta;
    w = sqrt((T)(S)(delta / sc)); // dL is kept in S precision by the reference (factor.hpp:166)
  }
  if (Jc) {
#pragma unroll
    for (int j = 0; j < 9; j++) {
      const T a0 = Jc[2 * j], a1 = Jc[2 * j + 1];
      Jc[2 * j] = w * (u00 * a0 + u01 * a1);
      Jc[2 * j + 1] = w * (u11 * a1);
    }
#pragma unroll
    for (int j = 0; j < 3; j++) {
      const T a0 = Jp[2 * j], a1 = Jp[2 * j + 1];
      Jp[2 * j] = w * (u00 * a0 + u01 * a1);
      Jp[2 * j + 1] = w * (u11 * a1);
    }
    r[0] = w * e0;
    r[1] = w * e1;
  }
  return chi;
}

// ---------------------------------------------------------------------------------------------
// K1: factor evaluation + point-side assembly + camera-side partials of diag(B) and g_c
//     replaces compute_error_kernel / compute_jacobian_kernel / compute_chi2_kernel /
//     compute_hessian_scalar_diagonal_kernel / compute_b_kernel (ops/error.hpp:252, ops/linearize.hpp:10,238,
//     ops/chi2.hpp:32, ops/hessian.hpp:418)
// ---------------------------------------------------------------------------------------------
// RESCALE = true is the second pass of a low-precision (bf16) linearisation: the reference casts the Jacobians to S when
// they are evaluated (ops/linearize.hpp:43-64), computes the Jacobi scales from those rounded values and then scales
// them IN PLACE, rounding to S a second time (scale_jacobians_kernel, ops/linearize.hpp:140-180:
// J = (S)((T)J * scale)).  The pass reads the stored (rounded, unscaled) Jacobians and residuals back, applies exactly
// that second rounding, stores J~ and assembles from J~; every later kernel then works in the scaled space with D = I.
template <typename T, typename S, bool EXT, bool RESCALE = false>
__global__ void __launch_bounds__(TILE, LIN_MIN_BLOCKS)
k_linearize(DevStruct ds, const T *__restrict__ cams, const T *__restrict__ pts,
            const typename V2<T>::type *__restrict__ obs, typename V2<S>::type *__restrict__ J,
            typename V2<T>::type *__restrict__ res, T *__restrict__ Cg, T *__restrict__ part /*[nrows][18]*/,
            double *__restrict__ cost_part /*[ntiles]*/, Robust rb, ExtFactor ex) {
  const bool robust = rb.Pu != nullptr || rb.loss_kind != 0;
  // ncu of the first version (tables read from global memory inside the reduction loops, two 9-wide camera passes):
  // 41 % of the stall samples were long-scoreboard waits in those loops, 61 % of all samples sat in the loops.
  // Here the tile's packed record (segment and point tables) is copied to shared memory while the camera model is
  // evaluated, and diag(B) and g_c go through ONE 18-wide staged pass with one thread per (segment, component triple).
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T *sv = reinterpret_cast<T *>(smem_raw);            // [TILE*LIN_STR] staging (point side uses the first TILE*9)
  T *acc = sv + TILE * LIN_STR;                       // [SLOT_CAP*18]
  unsigned char *rec = smem_raw + (((TILE * LIN_STR + SLOT_CAP * 18) * (int)sizeof(T) + 15) & ~15); // [REC_BYTES]
  // the super-tile's cameras (<= SLOT_CAP rows of the precomputed model terms) and the tile's points live in shared
  // memory: every observation then reads its 24 camera terms and 3 coordinates with LDS instead of a dependent chain of
  // global gathers (ncu of the gather version: long-scoreboard was the top stall, 4.7 cycles per issued instruction)
  T *cxs = reinterpret_cast<T *>(rec + REC_BYTES);    // [SLOT_CAP*LIN_CXS]
  T *pxs = cxs + SLOT_CAP * LIN_CXS;                  // [LIN_PTS]
  __shared__ double shd[32];
  const int st = blockIdx.x, t = threadIdx.x;
  const int row0 = ds.st_row[st], nslots = ds.st_row[st + 1] - row0;
  for (int i = t; i < nslots * 18; i += TILE) acc[i] = T(0);
  if (!EXT && !RESCALE) {
    for (int i = t; i < nslots * CAMX; i += TILE) {
      const int cs = i / CAMX, k = i - cs * CAMX;
      cxs[cs * LIN_CXS + k] = __ldg(cams + (int64_t)ds.row_cam[row0 + cs] * CAMX + k);
    }
    // the first tile's point coordinates (a tile owns a contiguous range of <= 128 points); the following tiles' are
    // fetched during the reduction phases of the tile before them
    const TileMeta tm0 = ds.tmeta[ds.st_tile[st]];
    const T *src = pts + 3 * (int64_t)tm0.p0;
    if (t < tm0.np * 3) pxs[t] = src[t];
    if (t + TILE < tm0.np * 3) pxs[t + TILE] = src[t + TILE];
  }
  __syncthreads();
  // the per-slot inputs of the NEXT tile (meta word, camera, observation) are fetched one tile ahead, so that a tile
  // starts with the camera / point gathers instead of a chain of dependent loads
  const int tile_end = ds.st_tile[st + 1];
  int64_t slot_n = (int64_t)ds.st_tile[st] * TILE + t;
  uint32_t om_n = ds.ometa[slot_n];
  int c_n = ds.tile_cam[slot_n];
  typename V2<T>::type ov_n = obs[slot_n];
  // ... and so are the tile's packed record (one 16-byte piece per thread, stored to shared memory at the top of its tile)
  // and its meta words.  (ncu source page of the version that loaded them at the top of the tile: the shared-memory
  // store of the record waited for its own global load - 6 % of all stall samples - and the first use of the tile meta
  // another 4 %, profiles/r2_ncu_stages_summary.txt.)
  const int tile_first = ds.st_tile[st];
  uint4 rec_n = make_uint4(0u, 0u, 0u, 0u);
  if (t < REC_BYTES / 16) rec_n = __ldg(reinterpret_cast<const uint4 *>(ds.trec + (int64_t)tile_first * REC_BYTES) + t);
  TileMeta tm_n = ds.tmeta[tile_first];
  for (int tile = tile_first; tile < tile_end; tile++) {
    // the record is consumed after the first barrier below
    if (t < REC_BYTES / 16) reinterpret_cast<uint4 *>(rec)[t] = rec_n;
    const TileMeta tm = tm_n;
    const int64_t slot = (int64_t)tile * TILE + t;
    const uint32_t om = om_n;
    const int c = c_n;
    const typename V2<T>::type ov = ov_n;
    if (tile + 1 < tile_end) {
      om_n = ds.ometa[slot + TILE];
      c_n = ds.tile_cam[slot + TILE];
      ov_n = obs[slot + TILE];
      if (t < REC_BYTES / 16) rec_n = __ldg(reinterpret_cast<const uint4 *>(ds.trec + (int64_t)(tile + 1) * REC_BYTES) + t);
      tm_n = ds.tmeta[tile + 1];
    }
    const int rank = (int)((om >> 8) & 0xffu), ptl = (int)(om & 0xffu);
    const bool active = t < tm.n;
    BalObs<T> B;
    double cost = 0.0;
    if (active) {
      if (RESCALE) {
        const int p = tm.p0 + ptl;
        load_J<T, S>(J, tile, t, B.Jc, B.Jp);
        const typename V2<T>::type rv = res[slot];
        B.r[0] = rv.x; B.r[1] = rv.y;
        const T *sc = reinterpret_cast<const T *>(ex.rescale) + (int64_t)c * 9;
        const T *sp = reinterpret_cast<const T *>(ex.rescale) + 9 * (int64_t)ds.Nc + 3 * (int64_t)p;
#pragma unroll
        for (int j = 0; j < 9; j++) {
          B.Jc[2 * j] = (T)(S)(B.Jc[2 * j] * sc[j]);
          B.Jc[2 * j + 1] = (T)(S)(B.Jc[2 * j + 1] * sc[j]);
        }
#pragma unroll
        for (int j = 0; j < 3; j++) {
          B.Jp[2 * j] = (T)(S)(B.Jp[2 * j] * sp[j]);
          B.Jp[2 * j + 1] = (T)(S)(B.Jp[2 * j + 1] * sp[j]);
        }
      } else if (EXT) {
        // the caller's kernel has evaluated this factor: fetch its residual and Jacobians
        const int64_t u = ex.slot_src[slot];
        const T *er = reinterpret_cast<const T *>(ex.r) + 2 * u;
        const T *ec = reinterpret_cast<const T *>(ex.Jc) + 18 * u;
        const T *ep = reinterpret_cast<const T *>(ex.Jp) + 6 * u;
        B.r[0] = er[0]; B.r[1] = er[1];
#pragma unroll
        for (int j = 0; j < 18; j++) B.Jc[j] = ec[j];
#pragma unroll
        for (int j = 0; j < 6; j++) B.Jp[j] = ep[j];
      } else {
        T X[3], ob[2];
        const T *cx = cxs + (int)(om >> 16) * LIN_CXS; // `cams` is the per-camera precomputed table (k_cam_precompute)
        X[0] = pxs[3 * ptl];
        X[1] = pxs[3 * ptl + 1];
        X[2] = pxs[3 * ptl + 2];
        ob[0] = ov.x;
        ob[1] = ov.y;
        bal_residual_jacobian_pre<T>(cx, X, ob, B);
      }
      res[slot] = V2<T>::make(B.r[0], B.r[1]); // the residual itself (export); the assembly below uses the whitened one
      if (robust) cost = (double)whiten_factor<T, S>(rb, slot, B.r, B.Jc, B.Jp);
      else cost = (double)(B.r[0] * B.r[0] + B.r[1] * B.r[1]);
      typename V2<S>::type *base = J + ((int64_t)tile * NPLANES) * TILE + t;
#pragma unroll
      for (int j = 0; j < 9; j++) base[j * TILE] = V2<S>::make((S)B.Jc[2 * j], (S)B.Jc[2 * j + 1]);
#pragma unroll
      for (int j = 0; j < 3; j++) base[(9 + j) * TILE] = V2<S>::make((S)B.Jp[2 * j], (S)B.Jp[2 * j + 1]);
      // what is stored is what every later kernel reads: keep the assembly consistent with S
#pragma unroll
      for (int j = 0; j < 18; j++) B.Jc[j] = (T)(S)B.Jc[j];
#pragma unroll
      for (int j = 0; j < 6; j++) B.Jp[j] = (T)(S)B.Jp[j];
      // point side: C (6 unique) and g = -Jp^T r, staged at the observation's position in point order
      T *row = sv + rank * 9;
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
    // per-tile cost partial with the reduction tree of block_sum (k_cost_tiles): chi2 of linearize == chi2 of cost, bit
    // for bit.  Its warp partials ride on the barriers the staging needs anyway.
    {
      double cw = cost;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) cw += __shfl_down_sync(0xffffffffu, cw, o);
      if ((t & 31) == 0) shd[t >> 5] = cw;
    }
    __syncthreads();
    // every thread has its point in registers: the staged coordinates are dead, fetch the next tile's (stored before the
    // last barrier of this tile; the two reduction phases hide the latency)
    T px_n0 = T(0), px_n1 = T(0);
    if (!EXT && !RESCALE && tile + 1 < tile_end) {
      const int n3 = tm_n.np * 3;
      const T *src = pts + 3 * (int64_t)tm_n.p0;
      if (t < n3) px_n0 = src[t];
      if (t + TILE < n3) px_n1 = src[t + TILE];
    }
    if (t == 0) {
      double tot = 0.0;
      for (int i = 0; i < TILE / 32; i++) tot += shd[i];
      cost_part[tile] = tot;
    }
    {
      // one thread per (point, component triple)
      const uint16_t *pt = reinterpret_cast<const uint16_t *>(rec + REC_PT);
      for (int item = t; item < tm.np * 3; item += TILE) {
        const int q = item / 3, g = item - 3 * q;
        const int b = pt[q], e = pt[q + 1];
        T a0 = T(0), a1 = T(0), a2 = T(0);
        for (int rowi = b; rowi < e; rowi++) {
          const T *r = sv + rowi * 9 + 3 * g;
          a0 += r[0];
          a1 += r[1];
          a2 += r[2];
        }
        // (a fragment of a long track leaves its partial sums; k_frag_sum adds the fragments of the point)
        T *out = tm.frag ? reinterpret_cast<T *>(ds.frag_part) + (int64_t)((tm.frag & FRAG_MASK) - 1) * 9 + 3 * g
                         : Cg + (int64_t)(tm.p0 + q) * 9 + 3 * g;
        out[0] = a0;
        out[1] = a1;
        out[2] = a2;
      }
    }
    __syncthreads();
    // camera side: diag(B) (9) and g_c (9) of this observation at the slot's own row (slots are in camera order)
    {
      T *row = sv + t * LIN_STR;
#pragma unroll
      for (int k = 0; k < 9; k++) {
        row[k] = active ? B.Jc[2 * k] * B.Jc[2 * k] + B.Jc[2 * k + 1] * B.Jc[2 * k + 1] : T(0);
        row[9 + k] = active ? -(B.Jc[2 * k] * B.r[0] + B.Jc[2 * k + 1] * B.r[1]) : T(0);
      }
    }
    __syncthreads();
    {
      const uint32_t *sg = reinterpret_cast<const uint32_t *>(rec + REC_SEG);
      for (int item = t; item < tm.nseg * 6; item += TILE) {
        const int q = item / 6, g = item - 6 * q;
        const uint32_t e0 = sg[q], e1 = sg[q + 1];
        const int b = (int)(e0 >> 16), e = (int)(e1 >> 16), cs = (int)(e0 & 0xffffu);
        T a0 = T(0), a1 = T(0), a2 = T(0);
        for (int rowi = b; rowi < e; rowi++) {
          const T *r = sv + rowi * LIN_STR + 3 * g;
          a0 += r[0];
          a1 += r[1];
          a2 += r[2];
        }
        T *ar = acc + cs * 18 + 3 * g;
        ar[0] += a0;
        ar[1] += a1;
        ar[2] += a2;
      }
    }
    if (!EXT && !RESCALE && tile + 1 < tile_end) {
      pxs[t] = px_n0;
      if (t < LIN_PTS - TILE) pxs[t + TILE] = px_n1;
    }
    __syncthreads(); // this tile's staging reads are done: the next tile may write its record and staging
  }
  for (int i = t; i < nslots * 18; i += TILE) part[(int64_t)ds.row_out[row0 + i / 18] * 18 + i % 18] = acc[i];
}

// Camera side of linearize: diag(B), g_c, Jacobi scales s = 1/(eps + sqrt(diag)) (graph.hpp:262-270), b_c = s g_c.
template <typename T>
__global__ void __launch_bounds__(288)
k_cam_reduce_lin(DevStruct ds, const T *__restrict__ part, T *__restrict__ diagB, T *__restrict__ gc,
                 int do_finish, int scale_on, T *__restrict__ scale_c, T *__restrict__ b_c,
                 const unsigned char *__restrict__ fixed_c = nullptr) {
  __shared__ T sh[32 * 9];
  __shared__ T out[18];
  const int c = blockIdx.x;
  cam_gather<T>(ds, c, part, 18, 2, sh, out);
  if (threadIdx.x < 9) {
    const int k = threadIdx.x;
    diagB[c * 9 + k] = out[k];
    gc[c * 9 + k] = out[9 + k];
    if (do_finish) {
      // a FIXED vertex (vertex.hpp:254-266) keeps its slot with scale 0: J~ = 0, b = 0, step 0 - the same numbers for
      // every other variable as the reference gets by leaving its columns out of the system
      const T s = (fixed_c && fixed_c[c]) ? T(0) : (scale_on ? (T)(1.0 / (DBL_EPSILON + sqrt((double)out[k]))) : T(1));
      scale_c[c * 9 + k] = s;
      b_c[c * 9 + k] = s * out[9 + k];
    }
  }
}
// After a multi-GPU allreduce of diagB / gc.
template <typename T>
__global__ void k_cam_finish_lin(int n, int scale_on, const T *__restrict__ diagB, const T *__restrict__ gc,
                                 T *__restrict__ scale_c, T *__restrict__ b_c, const unsigned char *__restrict__ fixed_c = nullptr) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const T s = (fixed_c && fixed_c[i / 9]) ? T(0) : (scale_on ? (T)(1.0 / (DBL_EPSILON + sqrt((double)diagB[i]))) : T(1));
  scale_c[i] = s;
  b_c[i] = s * gc[i];
}

// dst[slot_of[i]] = src[perm ? perm[i] : i]   (observations: caller order -> tile-padded storage slots)
template <typename V>
__global__ void k_scatter_slots(int64_t n, const int32_t *__restrict__ slot_of, const int64_t *__restrict__ perm,
                                const V *__restrict__ src, V *__restrict__ dst) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[slot_of[i]] = src[perm ? perm[i] : i];
}

// Small results (scalars, PCG state) go to PINNED HOST memory by a one-warp kernel instead of a DMA copy: a device-to-
// host cudaMemcpyAsync queued behind a large host-to-device upload on another stream waited for that whole upload
// (measured: every LM iteration lost the 1.5 ms of the overlapped observation upload).
__global__ void k_store_host(uint32_t *__restrict__ host_dst, const uint32_t *__restrict__ src, int nwords) {
  for (int i = threadIdx.x; i < nwords; i += blockDim.x) host_dst[i] = src[i];
  __threadfence_system();
}

// device-to-device copy on the SMs (the restore of rejected steps must not queue behind copy-engine traffic)
template <typename V> __global__ void k_copy(int64_t n, const V *__restrict__ src, V *__restrict__ dst) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

// Up to three deterministic partial sums in one launch (block b sums array b): cost, rho of the points, rho of the
// cameras at the end of an LM trial step.  Same per-array order as k_sum_partials.
struct SumJob { const double *part; int n; int out_idx; };
__global__ void __launch_bounds__(1024) k_sum_partials3(SumJob j0, SumJob j1, SumJob j2, double *__restrict__ out) {
  __shared__ double shd[32];
  const SumJob j = blockIdx.x == 0 ? j0 : (blockIdx.x == 1 ? j1 : j2);
  double acc = 0.0;
  for (int i = threadIdx.x; i < j.n; i += blockDim.x) acc += j.part[i];
  const double tot = block_sum<double>(acc, shd);
  if (threadIdx.x == 0) out[j.out_idx] = tot;
}

// Deterministic sum of partials (single CTA).
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
                                int write_lin, const unsigned char *__restrict__ fixed_p = nullptr) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= Np) return;
  if (fixed_p && fixed_p[p]) { // fixed point: scale 0, b = 0, W = 0, h = 0 (it adds nothing to S, b_S and gets no step)
    if (write_lin) {
      for (int k = 0; k < 3; k++) { scale_p[3 * (int64_t)p + k] = T(0); b_p[3 * (int64_t)p + k] = T(0); }
    } else {
      for (int k = 0; k < 6; k++) W[(int64_t)p * WST<T>::value + k] = T(0);
      for (int k = 0; k < 3; k++) h[HST * (int64_t)p + k] = T(0);
    }
    return;
  }
  const T *cg = Cg + (int64_t)p * 9;
  const T c00 = cg[0], c01 = cg[1], c02 = cg[2], c11 = cg[3], c12 = cg[4], c22 = cg[5];
  const T g0 = cg[6], g1 = cg[7], g2 = cg[8];
  T s0 = T(1), s1 = T(1), s2 = T(1);
  if (scale_on) {
    s0 = (T)(1.0 / (DBL_EPSILON + sqrt((double)c00)));
    s1 = (T)(1.0 / (DBL_EPSILON + sqrt((double)c11)));
    s2 = (T)(1.0 / (DBL_EPSILON + sqrt((double)c22)));
  }
  if (write_lin) { // linearize: only the mu-independent part; W and h are computed by the call in prepare
    scale_p[3 * (int64_t)p] = s0; scale_p[3 * (int64_t)p + 1] = s1; scale_p[3 * (int64_t)p + 2] = s2;
    b_p[3 * (int64_t)p] = s0 * g0; b_p[3 * (int64_t)p + 1] = s1 * g1; b_p[3 * (int64_t)p + 2] = s2 * g2;
    return;
  }
  // scaled, damped C
  const T a00 = damp_value<T>(s0 * s0 * c00, mu, use_identity);
  const T a11 = damp_value<T>(s1 * s1 * c11, mu, use_identity);
  const T a22 = damp_value<T>(s2 * s2 * c22, mu, use_identity);
  const T a01 = s0 * s1 * c01, a02 = s0 * s2 * c02, a12 = s1 * s2 * c12;
  // inverse of the SPD block through its Cholesky factor A = L L^T, A^-1 = L^-T L^-1 (backward stable; the cofactor
  // formula used first lost ~cond(A) more digits on weakly damped, nearly degenerate points and showed up as a 5e-8
  // difference of the truncated-PCG step against the pivoted inverse of the reference, cublas matinvBatched)
  const T l00 = sqrt(a00);
  const T i00 = T(1) / l00;
  const T l10 = a01 * i00, l20 = a02 * i00;
  const T l11 = sqrt(a11 - l10 * l10);
  const T i11 = T(1) / l11;
  const T l21 = (a12 - l20 * l10) * i11;
  const T l22 = sqrt(a22 - l20 * l20 - l21 * l21);
  const T i22 = T(1) / l22;
  // M = L^-1 (lower): m00 = i00, m10 = -l10 i00 i11, m11 = i11, m20 = (l10 l21 i11 - l20) i00 i22, m21 = -l21 i11 i22, m22 = i22
  const T m10 = -l10 * i00 * i11, m20 = (l10 * l21 * i11 - l20) * i00 * i22, m21 = -l21 * i11 * i22;
  // A^-1 = M^T M
  const T v00 = i00 * i00 + m10 * m10 + m20 * m20, v01 = m10 * i11 + m20 * m21, v02 = m20 * i22;
  const T v11 = i11 * i11 + m21 * m21, v12 = m21 * i22, v22 = i22 * i22;
  const T w00 = s0 * s0 * v00, w01 = s0 * s1 * v01, w02 = s0 * s2 * v02;
  const T w11 = s1 * s1 * v11, w12 = s1 * s2 * v12, w22 = s2 * s2 * v22;
  T *w = W + (int64_t)p * WST<T>::value;
  w[0] = w00; w[1] = w01; w[2] = w02; w[3] = w11; w[4] = w12; w[5] = w22;
  h[HST * (int64_t)p] = w00 * g0 + w01 * g1 + w02 * g2;
  h[HST * (int64_t)p + 1] = w01 * g0 + w11 * g1 + w12 * g2;
  h[HST * (int64_t)p + 2] = w02 * g0 + w12 * g1 + w22 * g2;
}

// K3b: camera-side sums of the Schur diagonal blocks and of the reduced right-hand side
//   A_c = sum_o Jc^T (I - N_o) Jc   (N_o = Jp W Jp^T; equals B_c - sum E W E^T restricted to the diagonal)
//   u_c = sum_o Jc^T Jp h_p
// replaces execute_schur_multiplication on the diagonal pairs + execute_b_Schur_computation
// (schur.hpp:649-734, 901-920) and the block copy of block_jacobi_schur.hpp:126-137.
// (The first implementation, a tile kernel on the same TMA pipeline as the Schur product that pushed the 54 values
// per observation through shared-memory staging and segment sums, is described in profiles/README.md.)
// Camera-major form.  The tile kernel pushed 54 values per observation through shared-memory staging and
// segment sums (ncu: 220 M warp instructions, 17 % FP64 pipe, 20 % DRAM - instruction-bound in the reduction loops).
// Here one CTA owns a CHUNK of ONE camera's observations (camera-major index built on the host): every thread walks
// its observations, gathers the 24 Jacobian values from the tile-major store by slot (adjacent slots of one camera
// share sectors) plus W and h of the point, and keeps all 54 sums in registers - no staging, no segment tables; one
// shuffle tree per CTA at the end.  Output: one 54-value row per chunk, rows of a camera contiguous (cam_ch_ptr).
#ifndef PC_MIN_BLOCKS
#define PC_MIN_BLOCKS 3 // 166 registers, no spills: 314 us against 344 us at 2 (202 registers) and 436 us at 4 (spills)
#endif
constexpr int PC_THREADS = 128;
template <typename T, typename S>
__global__ void __launch_bounds__(PC_THREADS, PC_MIN_BLOCKS)
k_prepare_cams(DevStruct ds, const typename V2<S>::type *__restrict__ J, const T *__restrict__ W,
               const T *__restrict__ h, T *__restrict__ part /*[nchunks][54]*/) {
  using S2 = typename V2<S>::type;
  __shared__ T sm[PC_THREADS / 32][54];
  const int ch = blockIdx.x;
  const int b = ds.ch_ptr[ch], e = ds.ch_ptr[ch + 1];
  T acc[54];
#pragma unroll
  for (int v = 0; v < 54; v++) acc[v] = T(0);
  // the indices of the next observation are fetched one iteration ahead: the 17 gathers of an observation then issue
  // without waiting for a dependent index load
  int i = b + (int)threadIdx.x;
  int slot_n = 0, p_n = 0;
  if (i < e) { slot_n = ds.cm_slot[i]; p_n = ds.cm_pt[i]; }
  for (; i < e; i += PC_THREADS) {
    const int slot = slot_n, p = p_n;
    if (i + PC_THREADS < e) { slot_n = ds.cm_slot[i + PC_THREADS]; p_n = ds.cm_pt[i + PC_THREADS]; }
    const S2 *base = J + ((int64_t)(slot >> 8) * NPLANES) * TILE + (slot & (TILE - 1));
    T jc[18], jp[6];
#pragma unroll
    for (int j = 0; j < 9; j++) {
      const S2 v = __ldg(base + j * TILE);
      jc[2 * j] = (T)v.x;
      jc[2 * j + 1] = (T)v.y;
    }
#pragma unroll
    for (int j = 0; j < 3; j++) {
      const S2 v = __ldg(base + (9 + j) * TILE);
      jp[2 * j] = (T)v.x;
      jp[2 * j + 1] = (T)v.y;
    }
    const T *w = W + (int64_t)p * WST<T>::value;
    const T w00 = w[0], w01 = w[1], w02 = w[2], w11 = w[3], w12 = w[4], w22 = w[5];
    const T *hp = h + (int64_t)p * HST;
    const T h0 = hp[0], h1 = hp[1], h2 = hp[2];
    // rows of Jp: a = (jp[0], jp[2], jp[4]), b = (jp[1], jp[3], jp[5]);  N = Jp W Jp^T ; M = I - N
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
    const T q0 = jp[0] * h0 + jp[2] * h1 + jp[4] * h2;
    const T q1 = jp[1] * h0 + jp[3] * h1 + jp[5] * h2;
    int idx = 0;
#pragma unroll
    for (int a = 0; a < 9; a++) {
      // row a of A = Jc^T M Jc (upper part): (jc_a^T M) jc_j
      const T k0 = m00 * jc[2 * a] + m01 * jc[2 * a + 1], k1 = m01 * jc[2 * a] + m11 * jc[2 * a + 1];
#pragma unroll
      for (int j = a; j < 9; j++) {
        acc[idx] += k0 * jc[2 * j] + k1 * jc[2 * j + 1];
        idx++;
      }
    }
#pragma unroll
    for (int k = 0; k < 9; k++) acc[45 + k] += jc[2 * k] * q0 + jc[2 * k + 1] * q1;
  }
  // CTA total: shuffle tree inside each warp, then the warps in order
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int v = 0; v < 54; v++) {
    T a = acc[v];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_down_sync(0xffffffffu, a, o);
    if (lane == 0) sm[warp][v] = a;
  }
  __syncthreads();
  if (threadIdx.x < 54) {
    T tot = sm[0][threadIdx.x];
#pragma unroll
    for (int w2 = 1; w2 < PC_THREADS / 32; w2++) tot += sm[w2][threadIdx.x];
    part[(int64_t)ch * 54 + threadIdx.x] = tot;
  }
}

// Gauss-Jordan inverse of a symmetric positive definite 9x9 matrix by 162 threads on the augmented matrix
// M = [A | I] (9 x 18, row-major in shared memory).  S_cc is SPD with a unit-scale diagonal (Jacobi scaling plus
// damping), so no pivoting is needed.  Every thread of the block must call it (barriers inside).
template <typename T> __device__ __forceinline__ void invert9_block(T *M /*[9*18]*/, T *f /*[9]*/, int t) {
  const int i = t / 18, j = t - 18 * i; // element (i, j) for t < 162
  __syncthreads(); // the callers fill M right before the call (racecheck: the first pivot read raced with that fill)
  for (int c = 0; c < 9; c++) {
    const T piv = M[c * 18 + c];
    __syncthreads();
    if (t < 162) {
      if (i == c) M[t] = M[t] / piv;   // scale the pivot row
      if (j == c) f[i] = M[t];         // column c before elimination (f[c] is unused)
    }
    __syncthreads();
    if (t < 162 && i != c) M[t] -= f[i] * M[c * 18 + j];
    __syncthreads();
  }
}

// K3c: per camera — sum partial rows (or take the allreduced sums), scale, damp, invert.
//   S~_cc = D A D with diagonal + (damp(B~_kk) - B~_kk);  Minv = S~_cc^-1  (block_jacobi_schur.hpp:139-147)
//   b_S   = D (g_c - u_c)                                  (schur.hpp:901-920)
//   dterm = damp(B~_kk) - B~_kk  (added to S p on the diagonal)
template <typename T>
__global__ void __launch_bounds__(288)
k_cam_reduce_prepare(DevStruct ds, const T *__restrict__ part, int from_sums, T *__restrict__ sums /*[Nc][54]*/,
                     int finish, T mu, int use_identity, const T *__restrict__ diagB, const T *__restrict__ gc,
                     const T *__restrict__ scale_c, T *__restrict__ Sdiag /*[Nc][81]*/, T *__restrict__ Minv,
                     T *__restrict__ bS, T *__restrict__ dterm, P2P pp, int xchg) {
  __shared__ T sh[32 * 9];
  __shared__ T out[54];
  __shared__ T Maug[9 * 18], fcol[9];
  const int c = blockIdx.x, t = threadIdx.x;
  if (!from_sums) {
    cam_gather<T>(ds, c, part, 54, 6, sh, out, ds.cam_ch_ptr); // rows of k_prepare_cams: one per camera chunk
    if (xchg) {
      // multi-GPU over peer memory: this rank's 54 sums go straight into every rank's receive slot (p2p.cuh); the
      // second launch (from_sums) waits for the peers and adds the slots in rank order
      const unsigned long long epoch = p2p_next_epoch(pp);
      if (t < 54)
        for (int q = 0; q < pp.nranks; q++) p2p_slot<T>(pp, q, pp.rank, epoch)[(int64_t)c * 54 + t] = out[t];
      p2p_signal(pp, epoch, gridDim.x);
      return;
    }
    if (t < 54) sums[(int64_t)c * 54 + t] = out[t];
  } else {
    if (xchg) {
      const unsigned long long epoch = p2p_current_epoch(pp);
      p2p_wait(pp, epoch);
      if (t < 54) out[t] = p2p_sum<T>(pp, epoch, (long long)c * 54 + t);
    } else if (t < 54) {
      out[t] = sums[(int64_t)c * 54 + t];
    }
    __syncthreads();
  }
  if (!finish) return;
  if (t < 81) {
    const int i = t % 9, j = t / 9;
    const int a = i < j ? i : j, b = i < j ? j : i;
    const int idx = a * 9 - (a * (a - 1)) / 2 + (b - a); // packed upper index of (a,b), row-wise
    const T si = scale_c[c * 9 + i], sj = scale_c[c * 9 + j];
    const bool fx = scale_c[c * 9] == T(0); // fixed camera: identity block, zero right-hand side (its unknowns stay 0)
    T val = si * sj * out[idx];
    if (i == j) {
      const T bt = si * si * diagB[c * 9 + i];
      const T dt = fx ? T(0) : damp_value<T>(bt, mu, use_identity) - bt;
      val += dt;
      dterm[c * 9 + i] = dt;
      bS[c * 9 + i] = si * (gc[c * 9 + i] - out[45 + i]);
      if (fx) val = T(1);
    }
    Maug[i * 18 + j] = val;
    Maug[i * 18 + 9 + j] = (i == j) ? T(1) : T(0);
    Sdiag[(int64_t)c * 81 + i + 9 * j] = val;
  }
  invert9_block<T>(Maug, fcol, t);
  if (t < 81) {
    const int i = t % 9, j = t / 9;
    Minv[(int64_t)c * 81 + i + 9 * j] = Maug[i * 18 + 9 + j];
  }
}

// ---------------------------------------------------------------------------------------------
// K4: matrix-free Schur product; tiles streamed through a TMA pipeline.
//   xs = D_c x (10-padded rows).
//   y_o = Jc x_c ; t_p = sum_o Jp^T y_o ; w_p = W_p t_p ; z_o = Jp w_p ; v_o = Jc^T (y_o - z_o)
//   row[camera] += sum over the camera's segment of v_o      ( = (B - E W E^T) x restricted to the super-tile )
// replaces execute_schur_vector_multiply (schur.hpp:347-393) on an explicit S.
//
// A CTA has two WORKERS of 256 threads; worker w consumes the tiles of parity w of the CTA's tile sequence (ring index
// i) with its own accumulator rows (added in a fixed order at the end of a super-tile).  Shared memory:
//   - a J ring of 2 slots, one per worker.  The 48 KB Jacobian block of a tile is dead as soon as its values are in
//     registers, so right after the worker's first barrier the slot is refilled with the worker's NEXT tile (i + 2),
//     which then has the whole tile time to land;
//   - the small per-tile data that stays live during the tile (record 2.4 KB, W rows 6 KB) in its own 4-deep ring;
//   - staging areas per worker (the point staging aliases the camera staging, they are never live together).
// Tile i uses mbarrier i % 4; tiles i and i + 4 belong to the same worker, so a parity wait can never run a phase ahead.
// (History, profiles/README.md: a 3-stage ring refilled at the END of a tile polled its refill flag five times per
// tile: 249 us -> 232 us with the early refill.)
// FULL = true turns the same pipeline into the full-system product J^T J u of the matrix-free PCGSolver
// (solver/pcg.hpp:141-163, kernels compute_Jv / compute_JtPv in ops/product.hpp): the W stage then carries the point part
// of the direction (u_p, stride WST), y = Jc u_c + Jp u_p, the point sums sum_o Jp^T y go straight to out_p, and the
// camera rows accumulate Jc^T y.
// ---------------------------------------------------------------------------------------------
// NW = workers per CTA (256 threads each).  Two in general; three for (float, float), where the halved Jacobian slots and
// staging leave room and the kernel is bound by per-tile latency, not bytes.
template <typename T, typename S, int NW = 2> struct SchurSmem2 {
  static constexpr int J_BYTES = NPLANES * TILE * (int)sizeof(typename V2<S>::type);
  static constexpr int W_BYTES = TILE_PTS * WST<T>::value * (int)sizeof(T);
  static constexpr int META_BYTES = REC_BYTES + W_BYTES;
  static constexpr int NMETA = 2 * NW; // tiles i and i + NMETA belong to the same worker: a parity wait cannot run a phase ahead
  static constexpr int SV_BYTES = TILE * 9 * (int)sizeof(T);      // camera staging (point staging [TILE*3] aliases it)
  static constexpr int SW_BYTES = TILE_PTS * 3 * (int)sizeof(T);  // point sums
  static constexpr int STG_BYTES = SV_BYTES + SW_BYTES;           // per worker
  static constexpr int META_OFF = NW * J_BYTES;
  static constexpr int STG_OFF = META_OFF + NMETA * META_BYTES;
  static constexpr int XL_OFF = STG_OFF + NW * STG_BYTES;
  static constexpr int ACC_OFF = XL_OFF + SLOT_CAP * 9 * (int)sizeof(T);
  static constexpr int BAR_OFF = ACC_OFF + NW * SLOT_CAP * 9 * (int)sizeof(T);
  static constexpr int TOTAL = BAR_OFF + 128; // mbarriers: NMETA tile slots + the record buffers of k_pcg_solve
  static_assert(META_BYTES % 16 == 0 && STG_BYTES % 16 == 0, "TMA destinations must stay 16-byte aligned");
};

// One tile of the product, executed by the 256 threads of one worker after the tile's mbarrier wait.
//   Js / rec / Ws: the tile's J slot, packed record and W rows in shared memory; xl: camera vector rows of the
//   super-tile; acc: the worker's accumulator rows; sv / sw: the worker's staging.
//   refill(next_p0, next_np) is called by the worker's thread 0 right after the first barrier (the J slot is free):
//   it issues the bulk copies of the worker's next tile, whose point range comes with the record just consumed.
template <typename T, typename S, bool FULL, int NW = 2, typename Refill>
__device__ __forceinline__ void product_tile(int worker, int t, const typename V2<S>::type *Js, const unsigned char *rec,
                                             const T *Ws, const T *xl, T *acc, T *sv, T *sw, T *__restrict__ out_p,
                                             const DevStruct &ds, T *tk_cur /*[Np][3]: this PCG iteration's point sums are kept, or null*/,
                                             Refill &&refill) {
  using S2 = typename V2<S>::type;
  T *sv3 = sv; // point-order staging: dead before the camera staging is written
  const TileMeta tm = *reinterpret_cast<const TileMeta *>(rec + REC_META);
  const int next_p0 = reinterpret_cast<const int32_t *>(rec + REC_NEXT)[2 * (NW - 1)];  // tile + NW: the worker's next tile
  const int next_np = reinterpret_cast<const int32_t *>(rec + REC_NEXT)[2 * (NW - 1) + 1];
  const uint32_t om = reinterpret_cast<const uint32_t *>(rec + REC_OMETA)[t];
  const int cslot = (int)(om >> 16), rank = (int)((om >> 8) & 0xffu), ptl = (int)(om & 0xffu);
  T jc[18], jp[6], y0 = T(0), y1 = T(0);
#pragma unroll
  for (int j = 0; j < 9; j++) {
    const S2 v = Js[j * TILE + t];
    jc[2 * j] = (T)v.x;
    jc[2 * j + 1] = (T)v.y;
  }
#pragma unroll
  for (int j = 0; j < 3; j++) {
    const S2 v = Js[(9 + j) * TILE + t];
    jp[2 * j] = (T)v.x;
    jp[2 * j + 1] = (T)v.y;
  }
  {
    const T *x = xl + cslot * 9;
#pragma unroll
    for (int j = 0; j < 9; j++) {
      const T xv = x[j];
      y0 += jc[2 * j] * xv;
      y1 += jc[2 * j + 1] * xv;
    }
  }
  if (FULL) {
    const T *u = Ws + ptl * WST<T>::value;
    y0 += jp[0] * u[0] + jp[2] * u[1] + jp[4] * u[2];
    y1 += jp[1] * u[0] + jp[3] * u[1] + jp[5] * u[2];
  }
  sv3[rank * 3 + 0] = jp[0] * y0 + jp[1] * y1; // staged at the position in point order
  sv3[rank * 3 + 1] = jp[2] * y0 + jp[3] * y1;
  sv3[rank * 3 + 2] = jp[4] * y0 + jp[5] * y1;
  worker_sync(worker); // every thread of the worker has its J values in registers: the worker's J slot is free
  if (t == 0) refill(next_p0, next_np);
  {
    const uint16_t *pt = reinterpret_cast<const uint16_t *>(rec + REC_PT);
    for (int item = t; item < tm.np * 3; item += TILE) {
      const int q = item / 3, k = item - 3 * q;
      const int b = pt[q], e = pt[q + 1];
      T a = T(0);
      for (int row = b; row < e; row++) a += sv3[row * 3 + k];
      if (FULL) {
        // (fragment of a long track: partial sum, added to the point's by k_frag_sum after the launch)
        if (tm.frag) reinterpret_cast<T *>(ds.frag_part)[(int64_t)((tm.frag & FRAG_MASK) - 1) * 3 + k] = a;
        else out_p[(int64_t)(tm.p0 + q) * 3 + k] = a;
      } else {
        // (fragment of a long track: the sum over ALL observations of the point, formed before the product - k_frag_dots /
        // phase H of k_pcg_solve)
        const T tv = tm.frag ? __ldcg(reinterpret_cast<const T *>(ds.frag_t) + (int64_t)((tm.frag & FRAG_MASK) - 1) * 3 + k) : a;
        sw[item] = tv;
        // keep t_p(p_k) for the back-substitution: x = sum_k alpha_k p_k, so t_p(x) = sum_k alpha_k t_p(p_k) and the
        // back-substitution need not stream the Jacobians again (k_backsubst_points)
        if (tk_cur != nullptr && (!tm.frag || (tm.frag & FRAG_FIRST))) tk_cur[(int64_t)tm.p0 * 3 + item] = tv;
      }
    }
  }
  worker_sync(worker);
  {
    T d0 = y0, d1 = y1;
    if (!FULL) {
      const T t0 = sw[ptl * 3], t1 = sw[ptl * 3 + 1], t2 = sw[ptl * 3 + 2];
      const T *w = Ws + ptl * WST<T>::value;
      const T w0 = w[0] * t0 + w[1] * t1 + w[2] * t2;
      const T w1 = w[1] * t0 + w[3] * t1 + w[4] * t2;
      const T w2 = w[2] * t0 + w[4] * t1 + w[5] * t2;
      d0 = y0 - (jp[0] * w0 + jp[2] * w1 + jp[4] * w2);
      d1 = y1 - (jp[1] * w0 + jp[3] * w1 + jp[5] * w2);
    }
    // stage v = Jc^T d at the slot's own row: slots are in camera order, so camera segments are contiguous rows
#pragma unroll
    for (int k = 0; k < 9; k++) sv[t * 9 + k] = jc[2 * k] * d0 + jc[2 * k + 1] * d1;
  }
  worker_sync(worker);
  {
    // one thread per (segment, component triple): sum the segment's rows, add to the camera's accumulator row
    const uint32_t *sg = reinterpret_cast<const uint32_t *>(rec + REC_SEG);
    for (int item = t; item < tm.nseg * 3; item += TILE) {
      const int q = item / 3, g = item - 3 * q;
      const uint32_t e0 = sg[q], e1 = sg[q + 1];
      const int b = (int)(e0 >> 16), e = (int)(e1 >> 16), cs = (int)(e0 & 0xffffu);
      T a0 = T(0), a1 = T(0), a2 = T(0);
      for (int row = b; row < e; row++) {
        const T *r = sv + row * 9 + 3 * g;
        a0 += r[0];
        a1 += r[1];
        a2 += r[2];
      }
      T *ar = acc + cs * 9 + 3 * g;
      ar[0] += a0;
      ar[1] += a1;
      ar[2] += a2;
    }
  }
  worker_sync(worker); // staging and the meta slot are free for the worker's next tile
}

// bulk copies of one tile into ring position i: J slot i & 1 (= its worker), meta slot and mbarrier i & 3.  The J and
// record copies carry an L2 evict-first policy: the 1 GB stream is read once per pass, the ~70 MB of vectors every
// pass re-reads then stay in the 126 MB L2 (-4.7 % on the kernel).
template <typename T, typename S, int NW = 2>
__device__ __forceinline__ void product_issue(unsigned char *smem, uint64_t *bars, const DevStruct &ds,
                                              const typename V2<S>::type *__restrict__ J, const T *__restrict__ W,
                                              int tile, int i, int p0, int np, uint64_t pol) {
  using SM = SchurSmem2<T, S, NW>;
  unsigned char *jdst = smem + (i % NW) * SM::J_BYTES;
  unsigned char *mdst = smem + SM::META_OFF + (i % SM::NMETA) * SM::META_BYTES;
  uint64_t *bar = &bars[i % SM::NMETA];
  const uint32_t wbytes = (uint32_t)(np * WST<T>::value * (int)sizeof(T));
  mbar_expect_tx(bar, (uint32_t)(SM::J_BYTES + REC_BYTES) + wbytes);
  bulk_g2s_hint(jdst, J + (int64_t)tile * NPLANES * TILE, SM::J_BYTES, bar, pol);
  bulk_g2s_hint(mdst, ds.trec + (int64_t)tile * REC_BYTES, REC_BYTES, bar, pol);
  bulk_g2s(mdst + REC_BYTES, W + (int64_t)p0 * WST<T>::value, wbytes, bar);
}

// One launch = one product: one CTA per super-tile (exports, the full-system solver, the NCCL fallback path and the
// stage timers; the Schur PCG itself runs k_pcg_solve, pcg_solve.cuh, which embeds the same pipeline).
template <typename T, typename S, bool FULL>
__global__ void __launch_bounds__(2 * TILE, 1)
k_schur_product2(DevStruct ds, const typename V2<S>::type *__restrict__ J, const T *__restrict__ W,
                 const T *__restrict__ xs, T *__restrict__ part /*[nrows][9]*/, const int *__restrict__ done_flag,
                 T *__restrict__ out_p /*FULL: [Np][3]*/) {
  using SM = SchurSmem2<T, S>;
  using S2 = typename V2<S>::type;
  extern __shared__ __align__(128) unsigned char smem[];
  if (done_flag && *done_flag) return; // PCG already stopped: nothing to do (uniform across the grid)
  const int worker = threadIdx.x >> 8, t = threadIdx.x & (TILE - 1);
  T *xl = reinterpret_cast<T *>(smem + SM::XL_OFF);
  T *acc_all = reinterpret_cast<T *>(smem + SM::ACC_OFF);
  T *acc = acc_all + worker * SLOT_CAP * 9;
  T *sv = reinterpret_cast<T *>(smem + SM::STG_OFF + worker * SM::STG_BYTES);
  T *sw = reinterpret_cast<T *>(smem + SM::STG_OFF + worker * SM::STG_BYTES + SM::SV_BYTES);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + SM::BAR_OFF);
  const int st_begin = ds.cta_st[blockIdx.x], st_end = ds.cta_st[blockIdx.x + 1];
  const int tile0 = ds.st_tile[st_begin], ntl = ds.st_tile[st_end] - tile0;
  const uint64_t pol = l2_policy_evict_first();

  if (threadIdx.x == 0) {
    for (int s = 0; s < SM::NMETA; s++) mbar_init(&bars[s], 1);
    mbar_fence_init();
    fence_proxy_async();
    for (int i = 0; i < 2 && i < ntl; i++) {
      const TileMeta tm = ds.tmeta[tile0 + i];
      product_issue<T, S>(smem, bars, ds, J, W, tile0 + i, i, tm.p0, tm.np, pol);
    }
  }

  for (int st = st_begin; st < st_end; st++) {
    const int row0 = ds.st_row[st], nslots = ds.st_row[st + 1] - row0;
    for (int i = threadIdx.x; i < nslots * 9; i += 2 * TILE) {
      const int s = i / 9, k = i - 9 * s;
      xl[i] = xs[(int64_t)ds.row_cam[row0 + s] * CAM_STRIDE + k];
    }
    for (int i = t; i < nslots * 9; i += TILE) acc[i] = T(0);
    __syncthreads(); // also: the mbarriers are initialised before any thread waits on them
    const int ib = ds.st_tile[st] - tile0, ie = ds.st_tile[st + 1] - tile0; // ring indices of this super-tile
    for (int i = ib + ((ib ^ worker) & 1); i < ie; i += 2) {                 // worker w takes ring indices of parity w
      mbar_wait(&bars[i & 3], (uint32_t)((i >> 2) & 1));
      const S2 *Js = reinterpret_cast<const S2 *>(smem + (i & 1) * SM::J_BYTES);
      const unsigned char *rec = smem + SM::META_OFF + (i & 3) * SM::META_BYTES;
      const T *Ws = reinterpret_cast<const T *>(rec + REC_BYTES);
      product_tile<T, S, FULL>(worker, t, Js, rec, Ws, xl, acc, sv, sw, out_p, ds, (T *)nullptr, [&](int next_p0, int next_np) {
        if (i + 2 < ntl) {
          fence_proxy_async();
          product_issue<T, S>(smem, bars, ds, J, W, tile0 + i + 2, i + 2, next_p0, next_np, pol);
        }
      });
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nslots * 9; i += 2 * TILE) {
      const int s = i / 9, k = i - 9 * s;
      part[(int64_t)ds.row_out[row0 + s] * 9 + k] = acc_all[i] + acc_all[SLOT_CAP * 9 + i];
    }
    __syncthreads(); // xl / acc are rewritten by the next super-tile
  }
}

// ---------------------------------------------------------------------------------------------
// Long tracks.  A point observed by more cameras than one tile holds is cut into fragment tiles (structure.hpp).  Sums
// over the point's observations then have two levels:
//   * sums a tile WRITES (C and g of k_linearize, the point part of the full-system product): every fragment leaves its
//     partial in frag_part, k_frag_sum adds the fragments of a point in ascending order;
//   * the sum a tile READS before it can go on (t_p = sum_o Jp^T Jc x_c of the Schur product and of the back-
//     substitution): heavy_point_dot (one warp per point) forms it over all observations of the point BEFORE the tile
//     kernels run (k_frag_dots, or phase H of k_pcg_solve) and leaves it in frag_t for every fragment of the point.
// Both are exact restatements of the single-tile sums (other summation order), deterministic, atomic-free.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void k_frag_sum(int nheavy, const int32_t *__restrict__ hv_pt, const int32_t *__restrict__ hv_ptr,
                           const T *__restrict__ frag_part, int width, T *__restrict__ out /*[Np][width]*/) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nheavy * width) return;
  const int hp = i / width, k = i - hp * width;
  T a = T(0);
  for (int f = hv_ptr[hp]; f < hv_ptr[hp + 1]; f++) a += frag_part[(int64_t)f * width + k];
  out[(int64_t)hv_pt[hp] * width + k] = a;
}
// t = sum over all observations of long-track point number hp of Jp^T (Jc x_c), x_c(j) = getx(c, j), by ONE WARP (all 32
// lanes must call it; no shared memory, no barrier): lanes stride over the point's observations - which are contiguous
// slots of its fragment tiles, so the twelve plane loads are coalesced - then a fixed shuffle tree; the result goes to
// frag_t of all the point's fragments.  One warp per point keeps many points in flight per SM (a real BAL set can have
// thousands of long tracks); a 2000-observation track costs its warp ~60 dependent rounds, which only that warp waits for.
// The sum runs in DOUBLE whatever T is: a landmark seen by hundreds of cameras is typically far away and weakly
// constrained in depth, so its W has a large eigenvalue (~1 / damping) that amplifies the rounding error of a 400-term
// float sum.
template <typename T, typename S, typename GetX>
__device__ __forceinline__ void heavy_point_dot(const DevStruct &ds, const typename V2<S>::type *__restrict__ J, int hp, GetX &&getx) {
  const int lane = threadIdx.x & 31;
  const int p = ds.hv_pt[hp];
  const int b = ds.pptr[p], e = ds.pptr[p + 1];
  double a0 = 0.0, a1 = 0.0, a2 = 0.0;
  for (int o = b + lane; o < e; o += 32) {
    const int slot = ds.slot_of_obs[o], c = ds.cam_idx[o];
    T jc[18], jp[6];
    double y0 = 0.0, y1 = 0.0;
    load_J<T, S>(J, slot >> 8, slot & (TILE - 1), jc, jp);
#pragma unroll
    for (int j = 0; j < 9; j++) {
      const double xv = (double)getx(c, j);
      y0 += (double)jc[2 * j] * xv;
      y1 += (double)jc[2 * j + 1] * xv;
    }
    a0 += (double)jp[0] * y0 + (double)jp[1] * y1;
    a1 += (double)jp[2] * y0 + (double)jp[3] * y1;
    a2 += (double)jp[4] * y0 + (double)jp[5] * y1;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a0 += __shfl_down_sync(0xffffffffu, a0, o);
    a1 += __shfl_down_sync(0xffffffffu, a1, o);
    a2 += __shfl_down_sync(0xffffffffu, a2, o);
  }
  if (lane == 0) {
    T *ft = reinterpret_cast<T *>(ds.frag_t);
    for (int f = ds.hv_ptr[hp]; f < ds.hv_ptr[hp + 1]; f++) {
      ft[(int64_t)f * 3] = (T)a0;
      ft[(int64_t)f * 3 + 1] = (T)a1;
      ft[(int64_t)f * 3 + 2] = (T)a2;
    }
  }
}
// one warp per long-track point; xs = D_c x in 10-padded camera rows
constexpr int FRAG_DOT_WARPS = 8;
template <typename T, typename S>
__global__ void __launch_bounds__(FRAG_DOT_WARPS * 32)
k_frag_dots(DevStruct ds, const typename V2<S>::type *__restrict__ J, const T *__restrict__ xs, const int *__restrict__ done_flag) {
  if (done_flag && *done_flag) return;
  const int hp = (int)blockIdx.x * FRAG_DOT_WARPS + ((int)threadIdx.x >> 5);
  if (hp < ds.nheavy) heavy_point_dot<T, S>(ds, J, hp, [&](int c, int j) { return xs[c * CAM_STRIDE + j]; });
}

// ---------------------------------------------------------------------------------------------
// K5a: back-substitution + point update, one CTA per tile (schur.hpp:279-302 + ops/update.hpp:9-31):
//   t_p = sum_o Jp^T Jc (D_c x~_c) ; x~_p = (h_p - W_p t_p) / s_p ; delta_p = x~_p s_p ; rho partial.
// ---------------------------------------------------------------------------------------------
template <typename T, typename S>
__global__ void __launch_bounds__(TILE)
k_backsubst_tiles(DevStruct ds, const typename V2<S>::type *__restrict__ J, const T *__restrict__ W,
                  const T *__restrict__ xs, const T *__restrict__ h, const T *__restrict__ scale_p,
                  const T *__restrict__ b_p, T mu, T *__restrict__ pts, T *__restrict__ pts_bak,
                  T *__restrict__ delta_p, double *__restrict__ rho_part /*[ntiles]*/, int apply,
                  const T *__restrict__ scale_apply /*scales of the update x += delta~ * s (= scale_p unless J is pre-scaled)*/) {
  __shared__ T sv3[TILE * 3];
  __shared__ double shd[32];
  const int tile = blockIdx.x, t = threadIdx.x;
  const unsigned char *rec = ds.trec + (int64_t)tile * REC_BYTES;
  const TileMeta tm = *reinterpret_cast<const TileMeta *>(rec + REC_META);
  const uint32_t om = reinterpret_cast<const uint32_t *>(rec + REC_OMETA)[t];
  const int c = ds.tile_cam[(int64_t)tile * TILE + t];
  T jc[18], jp[6], y0 = T(0), y1 = T(0);
  load_J<T, S>(J, tile, t, jc, jp);
  {
    T x[10];
    load_cam<T>(xs, c, x);
#pragma unroll
    for (int j = 0; j < 9; j++) {
      y0 += jc[2 * j] * x[j];
      y1 += jc[2 * j + 1] * x[j];
    }
  }
  const int prank = (int)((om >> 8) & 0xffu);
  sv3[prank * 3 + 0] = jp[0] * y0 + jp[1] * y1; // staged at the position in point order
  sv3[prank * 3 + 1] = jp[2] * y0 + jp[3] * y1;
  sv3[prank * 3 + 2] = jp[4] * y0 + jp[5] * y1;
  __syncthreads();
  double rho = 0.0;
  // (long track: its first fragment takes the sum over all the point's observations from k_frag_dots and updates the
  // point; the other fragments have nothing to do)
  if (t < tm.np && (!tm.frag || (tm.frag & FRAG_FIRST))) {
    const uint16_t *pt = reinterpret_cast<const uint16_t *>(rec + REC_PT);
    const int b = pt[t], e = pt[t + 1];
    T t0 = T(0), t1 = T(0), t2 = T(0);
    if (tm.frag) {
      const T *ft = reinterpret_cast<const T *>(ds.frag_t) + (int64_t)((tm.frag & FRAG_MASK) - 1) * 3;
      t0 = ft[0]; t1 = ft[1]; t2 = ft[2];
    } else {
      for (int row = b; row < e; row++) {
        t0 += sv3[row * 3];
        t1 += sv3[row * 3 + 1];
        t2 += sv3[row * 3 + 2];
      }
    }
    const int p = tm.p0 + t;
    const T *w = W + (int64_t)p * WST<T>::value;
    const T wv[3] = {w[0] * t0 + w[1] * t1 + w[2] * t2, w[1] * t0 + w[3] * t1 + w[4] * t2,
                     w[2] * t0 + w[4] * t1 + w[5] * t2};
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const int64_t i = 3 * (int64_t)p + k;
      const T s = scale_p[i];
      const T xt = s != T(0) ? (h[HST * (int64_t)p + k] - wv[k]) / s : T(0); // scaled-space step of the point (0: fixed)
      delta_p[i] = xt;
      rho += (double)(xt * (mu * xt + b_p[i]));
      if (apply) {
        const T old = pts[i];
        pts_bak[i] = old;
        pts[i] = old + xt * scale_apply[i];
      }
    }
  }
  const double tot = block_sum<double>(rho, shd);
  if (t == 0) rho_part[tile] = tot;
}

// K5a': the same back-substitution WITHOUT the Jacobians.  The PCG solve kernel left t_p(p_k) of every iteration k it
// executed (DevStruct::tk) and the step lengths alpha_k (0 for an iterate that was rejected or never applied); the solution
// is x = sum_k alpha_k p_k, so t_p(x) = sum_k alpha_k t_p(p_k): one thread per point reads k_max x 3 values instead of the
// kernel above streaming 24 Jacobian values per observation (Venice FP64: 1.16 GB -> 0.33 GB per step, 236 -> ~70 us).
template <typename T>
__global__ void __launch_bounds__(256)
k_backsubst_points(DevStruct ds, int kmax, const T *__restrict__ W, const T *__restrict__ h, const T *__restrict__ scale_p,
                   const T *__restrict__ b_p, T mu, T *__restrict__ pts, T *__restrict__ pts_bak, T *__restrict__ delta_p,
                   double *__restrict__ rho_part /*[gridDim.x]*/, int apply, const T *__restrict__ scale_apply) {
  __shared__ double shd[32];
  __shared__ T al[32];
  if (threadIdx.x < 32) al[threadIdx.x] = (int)threadIdx.x < kmax ? reinterpret_cast<const T *>(ds.tk_alpha)[threadIdx.x] : T(0);
  __syncthreads();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  double rho = 0.0;
  if (p < ds.Np) {
    T t0 = T(0), t1 = T(0), t2 = T(0);
    const T *tk = reinterpret_cast<const T *>(ds.tk) + 3 * (int64_t)p;
    for (int k = 0; k < kmax; k++) {
      const T a = al[k];
      if (a == T(0)) continue; // (uniform across the block)
      const T *tp = tk + (int64_t)k * ds.Np * 3;
      t0 += a * tp[0];
      t1 += a * tp[1];
      t2 += a * tp[2];
    }
    const T *w = W + (int64_t)p * WST<T>::value;
    const T wv[3] = {w[0] * t0 + w[1] * t1 + w[2] * t2, w[1] * t0 + w[3] * t1 + w[4] * t2,
                     w[2] * t0 + w[4] * t1 + w[5] * t2};
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const int64_t i = 3 * (int64_t)p + k;
      const T s = scale_p[i];
      const T xt = s != T(0) ? (h[HST * (int64_t)p + k] - wv[k]) / s : T(0); // scaled-space step of the point (0: fixed)
      delta_p[i] = xt;
      rho += (double)(xt * (mu * xt + b_p[i]));
      if (apply) {
        const T old = pts[i];
        pts_bak[i] = old;
        pts[i] = old + xt * scale_apply[i];
      }
    }
  }
  const double tot = block_sum<double>(rho, shd);
  if (threadIdx.x == 0) rho_part[blockIdx.x] = tot;
}

// ---------------------------------------------------------------------------------------------
// PCG on the reduced camera system (solver/pcg_schur.hpp:79-168).  Scalars stay on the device; the two dot
// products are per-camera partials summed by every CTA in the same fixed order, so all CTAs (and all ranks)
// take identical decisions without a broadcast.  State is ping-ponged: kernel k reads st[k], CTA 0 writes
// st[k+1].
// ---------------------------------------------------------------------------------------------
template <typename T> struct PcgState {
  T rz, rz0, alpha, beta, denom;
  int iter, done, reason, pad;
};

constexpr int PCG_CAMS = 32; // cameras per CTA (288 threads)

template <typename T> __device__ __forceinline__ T sum_all(const T *a, int n, T *sh) {
  T acc = T(0);
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += a[i];
  return block_sum<T>(acc, sh);
}

// Camera side of the product: Ap_raw = D_c * sum(partial rows).  With finish != 0 (single GPU) it also forms
// Ap = Ap_raw + dterm p and the per-camera partial of p.Ap.
// With push != 0 (multi-GPU over peer memory) the nine values go straight into every rank's receive slot of this rank
// instead of Ap_raw, and the last CTA publishes the exchange (p2p.cuh); the consumer is k_pcg_update / k_p2p_sum.
template <typename T>
__global__ void __launch_bounds__(288)
k_cam_reduce_spmv(DevStruct ds, const T *__restrict__ part, const T *__restrict__ scale_c, T *__restrict__ Ap_raw,
                  int finish, const T *__restrict__ dterm, const T *__restrict__ p, T *__restrict__ Ap,
                  T *__restrict__ dot_part, const int *__restrict__ done_flag, P2P pp, int push) {
  __shared__ T sh[32 * 9];
  __shared__ T out[9];
  if (done_flag && *done_flag) return;
  const int c = blockIdx.x;
  cam_gather<T>(ds, c, part, 9, 1, sh, out);
  if (push) {
    const unsigned long long epoch = p2p_next_epoch(pp);
    if (threadIdx.x < 9) {
      const T raw = scale_c[c * 9 + threadIdx.x] * out[threadIdx.x];
      for (int r = 0; r < pp.nranks; r++) p2p_slot<T>(pp, r, pp.rank, epoch)[c * 9 + threadIdx.x] = raw;
    }
    p2p_signal(pp, epoch, gridDim.x);
    return;
  }
  if (threadIdx.x == 0) {
    T d = T(0);
#pragma unroll
    for (int k = 0; k < 9; k++) {
      const T raw = scale_c[c * 9 + k] * out[k];
      Ap_raw[c * 9 + k] = raw;
      if (finish) {
        const T pk = p[c * 9 + k];
        const T ap = raw + dterm[c * 9 + k] * pk;
        Ap[c * 9 + k] = ap;
        d += pk * ap;
      }
    }
    if (finish) dot_part[c] = d;
  }
}
// Multi-GPU: after the allreduce of Ap_raw.
template <typename T>
__global__ void k_dot_partials(int Nc, const T *__restrict__ Ap_raw, const T *__restrict__ dterm,
                               const T *__restrict__ p, T *__restrict__ Ap, T *__restrict__ dot_part,
                               const int *__restrict__ done_flag) {
  if (done_flag && *done_flag) return;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= Nc) return;
  T d = T(0);
#pragma unroll
  for (int k = 0; k < 9; k++) {
    const T pk = p[c * 9 + k];
    const T ap = Ap_raw[c * 9 + k] + dterm[c * 9 + k] * pk;
    Ap[c * 9 + k] = ap;
    d += pk * ap;
  }
  dot_part[c] = d;
}

// x = 0 ; r = bS ; z = Minv r ; p = z ; xs = D p ; rz partial per camera.
template <typename T>
__global__ void __launch_bounds__(288)
k_pcg_init(int Nc, const T *__restrict__ bS, const T *__restrict__ Minv, const T *__restrict__ scale_c,
           T *__restrict__ x, T *__restrict__ r, T *__restrict__ z, T *__restrict__ p, T *__restrict__ xs,
           T *__restrict__ rz_part) {
  __shared__ T sr[288], sq[288];
  const int t = threadIdx.x, g = t / 9, c = blockIdx.x * PCG_CAMS + g, k = t - 9 * g;
  const bool ok = c < Nc;
  const int i = c * 9 + k;
  sr[t] = ok ? bS[i] : T(0);
  __syncthreads();
  T acc = T(0);
  if (ok) {
    const T *m = Minv + (int64_t)c * 81;
    const T *rc = sr + g * 9;
#pragma unroll
    for (int j = 0; j < 9; j++) acc += m[k + 9 * j] * rc[j];
    x[i] = T(0);
    r[i] = sr[t];
    z[i] = acc;
    p[i] = acc;
    xs[c * CAM_STRIDE + k] = scale_c[i] * acc;
    if (k == 0) xs[c * CAM_STRIDE + 9] = T(0);
  }
  sq[t] = sr[t] * acc;
  __syncthreads();
  if (ok && k == 0) {
    T d = T(0);
#pragma unroll
    for (int j = 0; j < 9; j++) d += sq[g * 9 + j];
    rz_part[c] = d;
  }
}
template <typename T>
__global__ void __launch_bounds__(1024)
k_pcg_init_state(int Nc, const T *__restrict__ rz_part, PcgState<T> *st, int *done_flag) {
  __shared__ T sh[32];
  const T rz = sum_all<T>(rz_part, Nc, sh);
  if (threadIdx.x == 0) {
    *done_flag = 0; // (a cudaMemsetAsync here would run on a copy engine and queue behind an overlapped upload)
    PcgState<T> s;
    s.rz = rz; s.rz0 = (T)INFINITY; s.alpha = T(0); s.beta = T(0); s.denom = T(0);
    s.iter = 0; s.done = 0; s.reason = 0; s.pad = 0;
    st[0] = s;
  }
}

// first half of an iteration (Ap and the p.Ap partials are ready): denom, alpha, x/r update, z = Minv r, r.z partials.
// Returns false when the iteration stops here (uniform across the grid: every CTA sees the same state and sums).
template <typename T>
__device__ __forceinline__ bool pcg_half1(int Nc, const PcgState<T> *sin, PcgState<T> *sout, const T *dot_part, const T *Ap,
                                          const T *Minv, T *x, T *xbak, T *r, T *z, const T *p, T *rz_part, int *done_flag) {
  __shared__ T sh[32];
  __shared__ T sr[288], sq[288];
  PcgState<T> s = *sin;
  const bool leader = blockIdx.x == 0 && threadIdx.x == 0;
  if (s.done) {
    if (leader) *sout = s;
    return false;
  }
  if (s.rz == T(0)) { // pcg_schur.hpp:109-111
    if (leader) { s.done = 1; s.reason = 3; *sout = s; *done_flag = 1; }
    return false;
  }
  const T denom = sum_all<T>(dot_part, Nc, sh);
  if (denom == T(0) || isnan(denom)) { // :120-122
    if (leader) { s.done = 1; s.reason = 4; s.denom = denom; *sout = s; *done_flag = 1; }
    return false;
  }
  const T alpha = s.rz / denom;
  const int t = threadIdx.x, g = t / 9, c = blockIdx.x * PCG_CAMS + g, k = t - 9 * g;
  const bool ok = c < Nc;
  const int i = c * 9 + k;
  T rn = T(0);
  if (ok) {
    const T pi = p[i];
    const T xo = x[i];
    xbak[i] = xo;
    x[i] = alpha * pi + xo;      // ops::axpy_async(x, alpha, p, x)
    rn = -alpha * Ap[i] + r[i];  // ops::axpy_async(r, -alpha, Ap, r)
    r[i] = rn;
  }
  sr[t] = rn;
  __syncthreads();
  T acc = T(0);
  if (ok) {
    const T *m = Minv + (int64_t)c * 81;
    const T *rc = sr + g * 9;
#pragma unroll
    for (int j = 0; j < 9; j++) acc += m[k + 9 * j] * rc[j];
    z[i] = acc;
  }
  sq[t] = rn * acc;
  __syncthreads();
  if (ok && k == 0) {
    T d = T(0);
#pragma unroll
    for (int j = 0; j < 9; j++) d += sq[g * 9 + j];
    rz_part[c] = d;
  }
  if (leader) {
    s.alpha = alpha;
    s.denom = denom;
    *sout = s;
  }
  return true;
}

// second half: rz_new, rejection / convergence tests, beta, p update, xs = D p
template <typename T>
__device__ __forceinline__ void pcg_half2(int Nc, const PcgState<T> *sin, PcgState<T> *sout, T tol, T ratio, int max_iter,
                                          const T *scale_c, const T *rz_part, T *x, const T *xbak, const T *z, T *p, T *xs,
                                          int *done_flag) {
  __shared__ T sh[32];
  PcgState<T> s = *sin;
  const bool leader = blockIdx.x == 0 && threadIdx.x == 0;
  if (s.done) {
    if (leader) *sout = s;
    return;
  }
  const T rzn = sum_all<T>(rz_part, Nc, sh);
  const int t = threadIdx.x, g = t / 9, c = blockIdx.x * PCG_CAMS + g, k = t - 9 * g;
  const bool ok = c < Nc;
  const int i = c * 9 + k;
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

// Both halves in one cooperative launch (all CTAs co-resident): the grid-wide sync replaces a kernel boundary.
// st[0] -> st[1] -> st[2] are consecutive entries of the ping-pong state array.
template <typename T>
__global__ void __launch_bounds__(288)
k_pcg_update(int Nc, PcgState<T> *st, T tol, T ratio, int max_iter, T *dot_part, T *Ap, const T *Minv, const T *scale_c,
             T *x, T *xbak, T *r, T *z, T *p, T *xs, T *rz_part, int *done_flag,
             const T *Ap_raw /*multi-GPU: all-reduced raw product, else null*/, const T *dterm, P2P pp, int pull) {
  if (Ap_raw != nullptr) {
    // multi-GPU: Ap = Ap_raw + dterm p and the per-camera p.Ap partials are formed here, after the all-reduce.
    // With pull != 0 the all-reduce itself happens here: wait for the peers' pushes of this iteration's exchange
    // and add the nranks slots in rank order (p2p.cuh).
    if (!st->done) {
      __shared__ T sq0[288];
      const int t = threadIdx.x, g = t / 9, c = blockIdx.x * PCG_CAMS + g, k = t - 9 * g;
      const bool ok = c < Nc;
      T prod = T(0);
      unsigned long long epoch = 0;
      if (pull) {
        epoch = p2p_current_epoch(pp);
        p2p_wait(pp, epoch);
      }
      if (ok) {
        const int i = c * 9 + k;
        const T pk = p[i];
        const T ap = (pull ? p2p_sum<T>(pp, epoch, i) : Ap_raw[i]) + dterm[i] * pk;
        Ap[i] = ap;
        prod = pk * ap;
      }
      sq0[t] = prod;
      __syncthreads();
      if (ok && k == 0) {
        T d = T(0);
#pragma unroll
        for (int j = 0; j < 9; j++) d += sq0[g * 9 + j];
        dot_part[c] = d;
      }
    }
    cooperative_groups::this_grid().sync();
  }
  const bool go = pcg_half1<T>(Nc, st, st + 1, dot_part, Ap, Minv, x, xbak, r, z, p, rz_part, done_flag);
  if (!go) { // uniform: no CTA reaches the grid sync; carry the stopped state forward
    if (blockIdx.x == 0 && threadIdx.x == 0) st[2] = st[1];
    return;
  }
  cooperative_groups::this_grid().sync();
  pcg_half2<T>(Nc, st + 1, st + 2, tol, ratio, max_iter, scale_c, rz_part, x, xbak, z, p, xs, done_flag);
}

// ---------------------------------------------------------------------------------------------
// helpers of the persistent PCG kernel (pcg_solve.cuh)
// ---------------------------------------------------------------------------------------------
// lanes 0..8 hold v (others 0): fixed-tree sum, valid in lane 0
template <typename T> __device__ __forceinline__ T sum9(T v) {
  v += __shfl_down_sync(0xffffffffu, v, 8);
  v += __shfl_down_sync(0xffffffffu, v, 4);
  v += __shfl_down_sync(0xffffffffu, v, 2);
  v += __shfl_down_sync(0xffffffffu, v, 1);
  return v;
}
// fixed-order sum of n values written by other CTAs of the running grid (L2 loads); valid in every thread
template <typename T> __device__ __forceinline__ T grid_total(const T *vals, int n, T *sh) {
  T acc = T(0);
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += __ldcg(vals + i);
  return block_sum<T>(acc, sh);
}

// two such sums with one pair of CTA barriers (sh: 64 values)
template <typename T>
__device__ __forceinline__ void grid_total2(const T *a, int na, const T *b, int nb, T *sh /*[64]*/, T &ta, T &tb) {
  T va = T(0), vb = T(0);
  for (int i = threadIdx.x; i < na; i += blockDim.x) va += __ldcg(a + i);
  for (int i = threadIdx.x; i < nb; i += blockDim.x) vb += __ldcg(b + i);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    va += __shfl_down_sync(0xffffffffu, va, o);
    vb += __shfl_down_sync(0xffffffffu, vb, o);
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (lane == 0) { sh[w] = va; sh[32 + w] = vb; }
  __syncthreads();
  ta = T(0); tb = T(0);
  for (int i = 0; i < nw; i++) { ta += sh[i]; tb += sh[32 + i]; }
}

// xs = D_c x (for the back-substitution) ; also camera update + rho partial (ops/update.hpp:9-31,
// levenberg_marquardt.hpp:34-41).  Single pass over the 9 Nc camera scalars.
template <typename T>
__global__ void k_cam_step(int n, const T *__restrict__ x, const T *__restrict__ scale_c, const T *__restrict__ b_c,
                           T mu, T *__restrict__ xs, T *__restrict__ cams, T *__restrict__ cams_bak,
                           T *__restrict__ delta_c, double *__restrict__ rho_part, int apply,
                           const T *__restrict__ scale_apply) {
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
      cams[c * CAM_STRIDE + k] = old + xt * scale_apply[i];
    }
  }
  const double tot = block_sum<double>(rho, shd);
  if (threadIdx.x == 0) rho_part[blockIdx.x] = tot;
}

// K5 cost: residual only (graph.hpp:221-234 compute_error + chi2); one CTA per tile
template <typename T, bool EXT>
__global__ void __launch_bounds__(TILE)
k_cost_tiles(DevStruct ds, const T *__restrict__ cams, const T *__restrict__ pts,
             const typename V2<T>::type *__restrict__ obs, double *__restrict__ cost_part /*[ntiles]*/, Robust rb,
             ExtFactor ex) {
  // COST_TILES consecutive tiles per CTA; the slot inputs of the next tile are fetched while the current one is
  // evaluated (one tile per CTA spent most of its time in the chain index -> camera / point gather -> arithmetic).
  // The per-tile partial and its reduction tree are unchanged: chi2 stays bit-identical to the linearize kernel's.
  __shared__ double shd[32];
  const int t = threadIdx.x;
  const int tile_begin = blockIdx.x * COST_TILES, tile_end = min(ds.ntiles, tile_begin + COST_TILES);
  int64_t slot_n = (int64_t)tile_begin * TILE + t;
  uint32_t om_n = ds.ometa[slot_n];
  int c_n = ds.tile_cam[slot_n];
  typename V2<T>::type ov_n = obs[slot_n];
  for (int tile = tile_begin; tile < tile_end; tile++) {
    const TileMeta tm = ds.tmeta[tile];
    const int64_t slot = (int64_t)tile * TILE + t;
    const uint32_t om = om_n;
    const int c = c_n;
    const typename V2<T>::type ov = ov_n;
    if (tile + 1 < tile_end) {
      om_n = ds.ometa[slot + TILE];
      c_n = ds.tile_cam[slot + TILE];
      ov_n = obs[slot + TILE];
    }
    double cost = 0.0;
    if (t < tm.n) {
      T r[2];
      if (EXT) {
        const T *er = reinterpret_cast<const T *>(ex.r) + 2 * (int64_t)ex.slot_src[slot];
        r[0] = er[0]; r[1] = er[1];
      } else {
        const int p = tm.p0 + (int)(om & 0xffu);
        T cx[CAMX], X[3], ob[2];
        load_camx<T>(cams, c, cx); // `cams` is the per-camera precomputed table (k_cam_precompute)
        X[0] = pts[3 * (int64_t)p];
        X[1] = pts[3 * (int64_t)p + 1];
        X[2] = pts[3 * (int64_t)p + 2];
        ob[0] = ov.x;
        ob[1] = ov.y;
        bal_residual_pre<T>(cx, X, ob, r);
      }
      if (rb.Pu != nullptr || rb.loss_kind != 0) cost = (double)whiten_factor<T, T>(rb, slot, r, (T *)nullptr, (T *)nullptr);
      else cost = (double)(r[0] * r[0] + r[1] * r[1]);
    }
    const double tot = block_sum<double>(cost, shd);
    if (t == 0) cost_part[tile] = tot;
  }
}

// ---------------------------------------------------------------------------------------------
// Full-system matrix-free PCG (solver/pcg.hpp:61-232 + preconditioner/block_jacobi.hpp): small vector kernels.
// The loop is DEVICE-RESIDENT: the PCG scalars (r.z, alpha, beta, 1/|r|, stop reason) live in a FullState on the device,
// written by one-CTA k_full_scalar launches from two-stage fixed-order reductions; every kernel of an iteration returns at
// once when the state says the solve has stopped, so the host enqueues whole batches of iterations without reading anything
// back (the reference: three blocking scalar reads per iteration, pcg.hpp:108-192).  Vector layout: [9 Nc camera scalars |
// 3 Np point scalars], scaled space.
// ---------------------------------------------------------------------------------------------
// PCG state of the full-system solver on the device
template <typename T> struct FullState {
  T rz, rz0, alpha, beta, sc /*1 / |r|*/, rzn;
  int iter, done, reason, pad;
};
// per-camera: full scaled block B~ (from the 45 packed sums), damped diagonal, inverse (block-parallel Gauss-Jordan)
template <typename T>
__global__ void __launch_bounds__(288)
k_full_cam_blocks(DevStruct ds, const T *__restrict__ part /*[nrows][54]*/, T mu, int use_identity,
                  const T *__restrict__ scale_c, T *__restrict__ Bfull /*[Nc][81] scaled, undamped*/, T *__restrict__ MinvF) {
  __shared__ T sh[32 * 9];
  __shared__ T out[54];
  __shared__ T Maug[9 * 18], fcol[9];
  const int c = blockIdx.x, t = threadIdx.x;
  cam_gather<T>(ds, c, part, 54, 5, sh, out, ds.cam_ch_ptr);
  if (t < 81) {
    const int i = t % 9, j = t / 9;
    const int a = i < j ? i : j, b = i < j ? j : i;
    const int idx = a * 9 - (a * (a - 1)) / 2 + (b - a);
    T val = scale_c[c * 9 + i] * scale_c[c * 9 + j] * out[idx];
    Bfull[(int64_t)c * 81 + i + 9 * j] = val;
    if (i == j) val = scale_c[c * 9] == T(0) ? T(1) : damp_value<T>(val, mu, use_identity); // fixed camera: identity
    Maug[i * 18 + j] = val;
    Maug[i * 18 + 9 + j] = (i == j) ? T(1) : T(0);
  }
  invert9_block<T>(Maug, fcol, t);
  if (t < 81) {
    const int i = t % 9, j = t / 9;
    MinvF[(int64_t)c * 81 + i + 9 * j] = Maug[i * 18 + 9 + j];
  }
}
// clamped scalar diagonal of J~^T J~ (pcg.hpp:93-104): cameras from diag(B), points from diag(C)
template <typename T>
__device__ __forceinline__ T full_diag(int64_t i, int64_t dimc, const T *diagB, const T *Cg, const T *scale) {
  T d;
  if (i < dimc) d = diagB[i];
  else {
    const int64_t q = (i - dimc) / 3;
    const int k = (int)((i - dimc) % 3);
    d = Cg[q * 9 + (k == 0 ? 0 : (k == 1 ? 3 : 5))];
  }
  d = scale[i] * scale[i] * d;
  return fmin(fmax(d, (T)1.0e-6), (T)1.0e32);
}
// v2 = [Ap_raw_c | D_p out_p] + mu (diag or 1) p ; per-block partial of p . v2
template <typename T>
__global__ void __launch_bounds__(256)
k_full_finish_v2(int Nc, int Np, T mu, int use_identity, const T *__restrict__ Ap_raw, const T *__restrict__ out_p,
                 const T *__restrict__ p, const T *__restrict__ scale, const T *__restrict__ diagB, const T *__restrict__ Cg,
                 T *__restrict__ v2, T *__restrict__ partial, const int *__restrict__ done_flag) {
  __shared__ T sh[32];
  if (done_flag && *done_flag) return;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t dimc = 9 * (int64_t)Nc, n = dimc + 3 * (int64_t)Np;
  T prod = T(0);
  if (i < n) {
    const T pv = p[i];
    T v = i < dimc ? Ap_raw[i] : scale[i] * out_p[i - dimc];
    v += use_identity ? mu * pv : mu * full_diag<T>(i, dimc, diagB, Cg, scale) * pv;
    v2[i] = v;
    prod = pv * v;
  }
  const T tot = block_sum<T>(prod, sh);
  if (threadIdx.x == 0) partial[blockIdx.x] = tot;
}
template <typename T>
__global__ void __launch_bounds__(256)
k_vec_dot(int64_t n, const T *__restrict__ a, const T *__restrict__ b, T *__restrict__ partial) {
  __shared__ T sh[32];
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const T tot = block_sum<T>(i < n ? a[i] * b[i] : T(0), sh);
  if (threadIdx.x == 0) partial[blockIdx.x] = tot;
}
// y = r / |r| ; z = M^-1 y: 9x9 blocks for the cameras (MinvF), 3x3 for the points (D^-1 W D^-1, W from k_point_prepare);
// per-block partial of r.z   (pcg.hpp:108-121, 184-192; y is formed on the fly with the rounding of the stored vector)
template <typename T>
__global__ void __launch_bounds__(256)
k_full_precond(int Nc, int Np, const FullState<T> *__restrict__ st, const T *__restrict__ MinvF, const T *__restrict__ W,
               const T *__restrict__ scale, const T *__restrict__ r, T *__restrict__ z, T *__restrict__ partial, int check_done) {
  __shared__ T sh[32];
  if (check_done && st->done) return;
  const T sc = st->sc;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t dimc = 9 * (int64_t)Nc, n = dimc + 3 * (int64_t)Np;
  T prod = T(0);
  if (i < n) {
    T zi;
    if (i < dimc) {
      const int64_t c = i / 9;
      const int k = (int)(i % 9);
      const T *m = MinvF + c * 81;
      T acc = T(0);
#pragma unroll
      for (int j = 0; j < 9; j++) acc += m[k + 9 * j] * (sc * r[c * 9 + j]);
      zi = acc;
    } else {
      const int64_t q = (i - dimc) / 3;
      const int k = (int)((i - dimc) % 3);
      const T *w = W + q * WST<T>::value;
      const T *sp = scale + dimc + 3 * q, *rr = r + dimc + 3 * q;
      if (sp[0] == T(0)) {
        zi = T(0); // fixed point
      } else {
        const T a0 = (sc * rr[0]) / sp[0], a1 = (sc * rr[1]) / sp[1], a2 = (sc * rr[2]) / sp[2];
        const T v = k == 0 ? w[0] * a0 + w[1] * a1 + w[2] * a2
                  : (k == 1 ? w[1] * a0 + w[3] * a1 + w[4] * a2 : w[2] * a0 + w[4] * a1 + w[5] * a2);
        zi = v / sp[k];
      }
    }
    z[i] = zi;
    prod = r[i] * zi;
  }
  const T tot = block_sum<T>(prod, sh);
  if (threadIdx.x == 0) partial[blockIdx.x] = tot;
}
// One CTA: tot = fixed-order sum of the per-block partials, then the scalar step `stage` of the loop (pcg.hpp:93-232):
//   0  1/|r| from r.r at the start       1  rz = r.z at the start, state reset, rz == 0 / max_iter == 0 stops
//   2  alpha = rz / p.v2                  3  1/|r| from r.r
//   4  rz_new: iteration count, rejection (reason 2), beta, convergence (1), iteration limit (0), rz == 0 (3)
template <typename T>
__global__ void __launch_bounds__(1024)
k_full_scalar(int stage, const T *__restrict__ partial, int n, FullState<T> *st, T tol, T ratio, int max_iter) {
  __shared__ T sh[32];
  if (stage >= 2 && st->done) return;
  const T tot = sum_all<T>(partial, n, sh);
  if (threadIdx.x != 0) return;
  FullState<T> s = *st;
  if (stage == 0 || stage == 3) {
    s.sc = (T)(1.0 / sqrt(tot)); // y = r / |r| (pcg.hpp:108-121): the division runs in double as on the host
  } else if (stage == 1) {
    s.rz = tot; s.rz0 = (T)INFINITY; s.alpha = T(0); s.beta = T(0); s.rzn = T(0);
    s.iter = 0; s.done = 0; s.reason = 0; s.pad = 0;
    if (max_iter <= 0) s.done = 1;
    else if (tot == T(0)) { s.done = 1; s.reason = 3; }
  } else if (stage == 2) {
    s.alpha = s.rz / tot;
  } else {
    s.iter += 1;
    s.rzn = tot;
    if (fabs(tot) > ratio * s.rz0 || isnan(tot)) { // rejected iterate: x is restored by k_full_restore
      s.done = 1; s.reason = 2;
    } else {
      s.rz0 = fmin(s.rz0, fabs(tot));
      s.beta = tot / s.rz;
      s.rz = tot;
      if (fabs(tot) < tol) { s.done = 1; s.reason = 1; }
      else if (s.iter >= max_iter) { s.done = 1; s.reason = 0; }
      else if (tot == T(0)) { s.done = 1; s.reason = 3; }
    }
  }
  *st = s;
}
// p = first ? z : beta p + z (ops::axpy), then u = D p: cameras into the 10-padded rows read by the product kernel,
// points into the W-strided stage buffer
template <typename T>
__global__ void k_full_direction(int Nc, int Np, int first, const FullState<T> *__restrict__ st, T *__restrict__ p,
                                 const T *__restrict__ z, const T *__restrict__ scale, T *__restrict__ xs, T *__restrict__ upw) {
  if (st->done) return;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t dimc = 9 * (int64_t)Nc, n = dimc + 3 * (int64_t)Np;
  if (i >= n) return;
  const T pn = first ? z[i] : st->beta * p[i] + z[i];
  p[i] = pn;
  const T v = scale[i] * pn;
  if (i < dimc) xs[(i / 9) * CAM_STRIDE + i % 9] = v;
  else upw[((i - dimc) / 3) * WST<T>::value + (i - dimc) % 3] = v;
}
// xbak = x ; x += alpha p ; r -= alpha v2 ; per-block partial of r.r
template <typename T>
__global__ void __launch_bounds__(256)
k_full_update_xr(int64_t n, const FullState<T> *__restrict__ st, const T *__restrict__ p, const T *__restrict__ v2,
                 T *__restrict__ x, T *__restrict__ xbak, T *__restrict__ r, T *__restrict__ partial) {
  __shared__ T sh[32];
  if (st->done) return;
  const T alpha = st->alpha;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  T rn = T(0);
  if (i < n) {
    const T xo = x[i];
    xbak[i] = xo;
    x[i] = alpha * p[i] + xo;
    rn = -alpha * v2[i] + r[i];
    r[i] = rn;
  }
  const T tot = block_sum<T>(rn * rn, sh);
  if (threadIdx.x == 0) partial[blockIdx.x] = tot;
}
// rejected iterate (pcg.hpp:194-199): x = xbak
template <typename T>
__global__ void k_full_restore(int64_t n, const FullState<T> *__restrict__ st, const T *__restrict__ xbak, T *__restrict__ x) {
  if (st->reason != 2) return;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) x[i] = xbak[i];
}

// point part of the step for the full-system solver: delta_p = x_p, rho partial, backup + update (ops/update.hpp:9-31)
template <typename T>
__global__ void __launch_bounds__(256)
k_full_point_step(int64_t n3, const T *__restrict__ xp, const T *__restrict__ scale_p, const T *__restrict__ b_p, T mu,
                  T *__restrict__ pts, T *__restrict__ pts_bak, T *__restrict__ delta_p, double *__restrict__ rho_part,
                  int apply, const T *__restrict__ scale_apply) {
  __shared__ double shd[32];
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double rho = 0.0;
  if (i < n3) {
    const T xt = xp[i];
    delta_p[i] = xt;
    rho = (double)(xt * (mu * xt + b_p[i]));
    if (apply) {
      const T old = pts[i];
      pts_bak[i] = old;
      pts[i] = old + xt * scale_apply[i];
    }
  }
  const double tot = block_sum<double>(rho, shd);
  if (threadIdx.x == 0) rho_part[blockIdx.x] = tot;
}

// ---------------------------------------------------------------------------------------------
// parity exports (test paths; not timed)
// ---------------------------------------------------------------------------------------------
// Scaled, undamped Hessian values in the reference layout (hessian.hpp:257-288):
// [B_0 .. B_{Nc-1}] then per point [E_{c1,p} E_{c2,p} ... C_p], blocks column-major.
// J~ is rounded to S after scaling (ops/linearize.hpp:176-178) as the reference does.
template <typename T, typename S>
__global__ void k_hessian_export(DevStruct ds, const typename V2<S>::type *__restrict__ J, const T *__restrict__ Cg,
                                 const T *__restrict__ scale_c, const T *__restrict__ scale_p, S *__restrict__ vals,
                                 double *__restrict__ Bacc /*[Nc][81], zeroed*/) {
  const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (o < ds.M) {
    T jc[18], jp[6];
    const int slot = ds.slot_of_obs[o];
    load_J<T, S>(J, slot / TILE, slot % TILE, jc, jp);
    const int c = ds.cam_idx[o], p = ds.pt_idx[o];
    T a[18], b[6];
    for (int i = 0; i < 9; i++) {
      a[2 * i] = (T)(S)(jc[2 * i] * scale_c[c * 9 + i]);
      a[2 * i + 1] = (T)(S)(jc[2 * i + 1] * scale_c[c * 9 + i]);
    }
    for (int j = 0; j < 3; j++) {
      b[2 * j] = (T)(S)(jp[2 * j] * scale_p[3 * (int64_t)p + j]);
      b[2 * j + 1] = (T)(S)(jp[2 * j + 1] * scale_p[3 * (int64_t)p + j]);
    }
    // offset: 81 Nc + 27 o + 9 p  (every earlier point contributes its E blocks and one C block)
    S *dst = vals + 81 * (int64_t)ds.Nc + 27 * o + 9 * (int64_t)p;
    for (int j = 0; j < 3; j++)
      for (int i = 0; i < 9; i++) dst[i + 9 * j] = (S)(a[2 * i] * b[2 * j] + a[2 * i + 1] * b[2 * j + 1]);
    for (int j = 0; j < 9; j++)
      for (int i = 0; i < 9; i++)
        atomicAdd(Bacc + (int64_t)c * 81 + i + 9 * j, (double)(a[2 * i] * a[2 * j] + a[2 * i + 1] * a[2 * j + 1]));
  }
  if (o < ds.Np) {
    const int64_t p = o;
    const T *cg = Cg + p * 9;
    const T s[3] = {scale_p[3 * p], scale_p[3 * p + 1], scale_p[3 * p + 2]};
    const T cf[9] = {cg[0], cg[1], cg[2], cg[1], cg[3], cg[4], cg[2], cg[4], cg[5]};
    S *dst = vals + 81 * (int64_t)ds.Nc + 27 * (int64_t)ds.pptr[p + 1] + 9 * p;
    for (int k = 0; k < 9; k++) dst[k] = (S)(s[k % 3] * s[k / 3] * cf[k]);
  }
}
template <typename S> __global__ void k_copy_B(int n, const double *__restrict__ Bacc, S *__restrict__ vals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) vals[i] = (S)Bacc[i];
}

} // namespace gb

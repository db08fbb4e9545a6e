// explicit_schur.cuh — the EXPLICIT form of the Schur complement as a solve mode (gb_pcg_options.schur_mode).
//
// Replaces SchurComplement::update_values (schur.hpp:227-235: execute_Hpp_copy, execute_schur_multiplication with
// schur_block_product_kernel_dim_b, ops/schur.hpp:154-188, 81 atomics per (point, camera pair)) and
// execute_schur_vector_multiply (schur.hpp:347-393: block matvecs with atomics, upper blocks then transposes).
//   k_schur_build          S_ij = - D_i (sum_p E_ip W_p E_jp^T) D_j for every off-diagonal block of the upper block-CSC:
//                          one warp per block, its (point, pair) tuples - sorted by block once, on the host
//                          (structure.hpp: ExplicitSchur) - summed in a fixed order.  No atomics: bit-reproducible.
//                          The diagonal blocks are the production path's S_cc (k_cam_reduce_prepare).
//   k_pcg_solve_explicit   the whole PCG solve (pcg_schur.hpp:79-168) in one cooperative launch on the stored S: per
//                          iteration  p = beta p + z | barrier | rows of S p (a CTA per camera row, its warps share the
//                          row's blocks; the symmetric matrix is read through a row view: blocks as stored or
//                          transposed), multi-GPU: LL exchange of the partial rows, p.Ap | barrier | x, r, z, r.z |
//                          barrier.  S stays in the 126 MB L2 between iterations when it fits (41 MB at Dubrovnik).
// When is this form faster than the matrix-free one?  DESIGN.md section 3 has the measured rule.
#pragma once
#include "kernels.cuh"

namespace gb {

struct ExplicitDev {
  int32_t Nc, nblocks;
  const int64_t *tptr;                        // [nblocks + 1]
  const int32_t *tup_a, *tup_b, *tup_p;       // [ntuples]
  const int32_t *blk_row, *blk_col;           // [nblocks]
  const int32_t *diag_block;                  // [Nc]
  const int32_t *row_ptr, *row_ent, *row_other;
};

// One warp per block.  Lane l owns the entries e = l, l + 32, l + 64 (< 81) of the column-major 9x9 block.
template <typename T, typename S>
__global__ void __launch_bounds__(256)
k_schur_build(ExplicitDev ed, const typename V2<S>::type *__restrict__ J, const T *__restrict__ W,
              const T *__restrict__ scale_c, const T *__restrict__ Sdiag /*[Nc][81] damped diagonal blocks*/,
              T *__restrict__ values /*[nblocks][81]*/) {
  using S2 = typename V2<S>::type;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (b >= ed.nblocks) return;
  const int ca = ed.blk_row[b], cb = ed.blk_col[b];
  T *out = values + (int64_t)b * 81;
  if (ca == cb) { // diagonal block: already reduced, scaled and damped by the production path
    for (int e = lane; e < 81; e += 32) out[e] = Sdiag[(int64_t)ca * 81 + e];
    return;
  }
  T acc[3] = {T(0), T(0), T(0)};
  const int64_t t0 = ed.tptr[b], t1 = ed.tptr[b + 1];
  for (int64_t k = t0; k < t1; k++) {
    const int sa = ed.tup_a[k], sb = ed.tup_b[k], p = ed.tup_p[k];
    const S2 *ja = J + ((int64_t)(sa >> 8) * NPLANES) * TILE + (sa & (TILE - 1));
    const S2 *jb = J + ((int64_t)(sb >> 8) * NPLANES) * TILE + (sb & (TILE - 1));
    const T *w = W + (int64_t)p * WST<T>::value;
    const T w00 = w[0], w01 = w[1], w02 = w[2], w11 = w[3], w12 = w[4], w22 = w[5];
    T pa[6], pb[6];
#pragma unroll
    for (int j = 0; j < 3; j++) {
      const S2 va = __ldg(ja + (9 + j) * TILE), vb = __ldg(jb + (9 + j) * TILE);
      pa[2 * j] = (T)va.x; pa[2 * j + 1] = (T)va.y;
      pb[2 * j] = (T)vb.x; pb[2 * j + 1] = (T)vb.y;
    }
    // q_v = W Jp_b^T e_v (v = 0, 1);  N[u][v] = Jp_a[u] . q_v  (2x2)
    const T q00 = w00 * pb[0] + w01 * pb[2] + w02 * pb[4], q01 = w01 * pb[0] + w11 * pb[2] + w12 * pb[4],
            q02 = w02 * pb[0] + w12 * pb[2] + w22 * pb[4];
    const T q10 = w00 * pb[1] + w01 * pb[3] + w02 * pb[5], q11 = w01 * pb[1] + w11 * pb[3] + w12 * pb[5],
            q12 = w02 * pb[1] + w12 * pb[3] + w22 * pb[5];
    const T n00 = pa[0] * q00 + pa[2] * q01 + pa[4] * q02, n01 = pa[0] * q10 + pa[2] * q11 + pa[4] * q12;
    const T n10 = pa[1] * q00 + pa[3] * q01 + pa[5] * q02, n11 = pa[1] * q10 + pa[3] * q11 + pa[5] * q12;
#pragma unroll
    for (int u = 0; u < 3; u++) {
      const int e = lane + 32 * u;
      if (e < 81) {
        const int r = e % 9, c = e / 9;
        const S2 var = __ldg(ja + r * TILE), vbc = __ldg(jb + c * TILE);
        const T a0 = (T)var.x, a1 = (T)var.y, b0 = (T)vbc.x, b1 = (T)vbc.y;
        acc[u] += a0 * (n00 * b0 + n01 * b1) + a1 * (n10 * b0 + n11 * b1);
      }
    }
  }
#pragma unroll
  for (int u = 0; u < 3; u++) {
    const int e = lane + 32 * u;
    if (e < 81) out[e] = -scale_c[ca * 9 + e % 9] * scale_c[cb * 9 + e / 9] * acc[u];
  }
}

constexpr int XS_THREADS = 256, XS_WARPS = XS_THREADS / 32;

// The PCG solve on the stored S.  grid <= co-resident CTAs; camera rows are dealt to the CTAs round-robin.
template <typename T>
__global__ void __launch_bounds__(XS_THREADS)
k_pcg_solve_explicit(ExplicitDev ed, const T *__restrict__ Svals /*[nblocks][81]*/, const T *__restrict__ Minv,
                     const T *__restrict__ bS, T *x, T *xbak, T *r, T *z, T *p, T *Ap, T *cta_red /*[2][grid]*/,
                     PcgState<T> *st_out, T tol, T ratio, int max_iter, P2P pp, int multi) {
  namespace cg = cooperative_groups;
  cg::grid_group grid = cg::this_grid();
  __shared__ T red[64];
  __shared__ T ypart[XS_WARPS][9];
  __shared__ T rvec[9];
  const int G = gridDim.x, Nc = ed.Nc;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool leader = blockIdx.x == 0 && threadIdx.x == 0;
  const int rr = lane % 9, g = lane / 9; // lanes < 27: row rr of the block, column group g (columns 3g .. 3g+2)
  const unsigned long long epoch0 = multi ? p2p_current_epoch(pp) : 0ull;

  // per-CTA partial of a dot product -> cta_red[which * G + blockIdx.x]
  auto publish = [&](T v, int which) {
    v = block_sum<T>(v, red);
    if (threadIdx.x == 0) *(volatile T *)(cta_red + which * G + blockIdx.x) = v;
  };

  // ---- start: x = 0, r = b_S, z = M^-1 r, p = 0 (beta = 0 makes the first direction z), rz = r.z ----------------------
  T part_rz = T(0);
  for (int c = blockIdx.x; c < Nc; c += G) {
    if (threadIdx.x < 9) rvec[threadIdx.x] = bS[c * 9 + threadIdx.x];
    __syncthreads();
    if (threadIdx.x < 9) {
      const int k = threadIdx.x, i = c * 9 + k;
      const T *m = Minv + (int64_t)c * 81;
      T a = T(0);
#pragma unroll
      for (int j = 0; j < 9; j++) a += m[k + 9 * j] * rvec[j];
      x[i] = T(0); r[i] = rvec[k]; z[i] = a; p[i] = T(0);
      part_rz += rvec[k] * a;
    }
    __syncthreads();
  }
  publish(part_rz, 1);
  __threadfence();
  grid.sync();
  PcgState<T> s;
  s.rz = grid_total<T>(cta_red + G, G, red);
  s.rz0 = (T)INFINITY; s.alpha = T(0); s.beta = T(0); s.denom = T(0);
  s.iter = 0; s.done = 0; s.reason = 0; s.pad = 0;
  T beta = T(0);
  int k = 0;
  for (; k < max_iter; k++) {
    if (s.rz == T(0)) { s.done = 1; s.reason = 3; break; } // pcg_schur.hpp:109-111
    // ---- p = beta p + z  (every entry by one thread of the grid) ------------------------------------------------------
    for (int i = blockIdx.x * XS_THREADS + threadIdx.x; i < 9 * Nc; i += G * XS_THREADS) p[i] = fma(beta, p[i], __ldcg(z + i));
    __threadfence();
    grid.sync();
    // ---- Ap = S p: one CTA per camera row; its warps share the row's off-diagonal blocks, warp 0 adds the diagonal ----
    const unsigned long long epoch = epoch0 + (unsigned long long)k + 1ull;
    const unsigned int e32 = (unsigned int)epoch;
    T part_dot = T(0);
    for (int c = blockIdx.x; c < Nc; c += G) {
      T y = T(0);
      if (lane < 27) {
        const int e0 = ed.row_ptr[c], e1 = ed.row_ptr[c + 1];
        for (int q = e0 + warp; q < e1; q += XS_WARPS) {
          const int ent = ed.row_ent[q], other = ed.row_other[q];
          const T *blk = Svals + (int64_t)(ent >> 1) * 81;
          const T *pj = p + other * 9 + 3 * g; // written by other CTAs in this kernel: L2 loads
          const T p0 = __ldcg(pj), p1 = __ldcg(pj + 1), p2 = __ldcg(pj + 2);
          if (ent & 1) // transposed: (S^T)[rr][c] = S[c][rr], column-major element c + 9 rr
            y += blk[3 * g + 9 * rr] * p0 + blk[3 * g + 1 + 9 * rr] * p1 + blk[3 * g + 2 + 9 * rr] * p2;
          else
            y += blk[rr + 9 * (3 * g)] * p0 + blk[rr + 9 * (3 * g + 1)] * p1 + blk[rr + 9 * (3 * g + 2)] * p2;
        }
      }
      // the three column groups of a warp, then the warps in order
      const T y1 = __shfl_sync(0xffffffffu, y, rr + 9), y2 = __shfl_sync(0xffffffffu, y, rr + 18);
      if (lane < 9) ypart[warp][lane] = (y + y1) + y2;
      __syncthreads();
      if (threadIdx.x < 9) {
        const int kk = threadIdx.x, i = c * 9 + kk;
        T raw = T(0);
#pragma unroll
        for (int w = 0; w < XS_WARPS; w++) raw += ypart[w][kk];
        if (multi) { // this rank's partial row (its points only) to every rank, summed in rank order (p2p.cuh)
          for (int q = 0; q < pp.nranks; q++) ll_store(ll_slot(pp, q, pp.rank, epoch), (long long)i, raw, e32);
          raw = ll_sum<T>(pp, epoch, (long long)i, e32);
        }
        // diagonal block (already summed over the ranks, damped)
        const T *d = Svals + (int64_t)ed.diag_block[c] * 81;
        T dd = T(0);
#pragma unroll
        for (int j = 0; j < 9; j++) dd += d[kk + 9 * j] * __ldcg(p + c * 9 + j);
        const T ap = raw + dd;
        Ap[i] = ap;
        part_dot += __ldcg(p + i) * ap;
      }
      __syncthreads();
    }
    publish(part_dot, 0);
    __threadfence();
    grid.sync();
    const T denom = grid_total<T>(cta_red, G, red);
    if (denom == T(0) || isnan(denom)) { s.done = 1; s.reason = 4; s.denom = denom; k++; break; } // pcg_schur.hpp:120-122
    const T alpha = s.rz / denom;
    // ---- x += alpha p ; r -= alpha Ap ; z = M^-1 r ; r.z -------------------------------------------------------------
    T prz = T(0);
    for (int c = blockIdx.x; c < Nc; c += G) {
      if (threadIdx.x < 9) {
        const int i = c * 9 + threadIdx.x;
        const T xo = x[i];
        xbak[i] = xo;
        x[i] = alpha * __ldcg(p + i) + xo;
        const T rn = -alpha * Ap[i] + r[i];
        r[i] = rn;
        rvec[threadIdx.x] = rn;
      }
      __syncthreads();
      if (threadIdx.x < 9) {
        const int kk = threadIdx.x;
        const T *m = Minv + (int64_t)c * 81;
        T a = T(0);
#pragma unroll
        for (int j = 0; j < 9; j++) a += m[kk + 9 * j] * rvec[j];
        z[c * 9 + kk] = a;
        prz += rvec[kk] * a;
      }
      __syncthreads();
    }
    publish(prz, 1);
    __threadfence();
    grid.sync();
    const T rzn = grid_total<T>(cta_red + G, G, red);
    s.iter += 1;
    s.alpha = alpha;
    s.denom = denom;
    if (fabs(rzn) > ratio * s.rz0 || isnan(rzn)) { // pcg_schur.hpp:144-148
      for (int i = blockIdx.x * XS_THREADS + threadIdx.x; i < 9 * Nc; i += G * XS_THREADS) x[i] = __ldcg(xbak + i);
      s.done = 1; s.reason = 2; s.rz = rzn;
      k++;
      break;
    }
    s.rz0 = fmin(s.rz0, fabs(rzn));
    beta = rzn / s.rz;
    s.beta = beta;
    s.rz = rzn;
    if (fabs(rzn) < tol) { s.done = 1; s.reason = 1; k++; break; }
  }
  if (!s.done) { s.done = 1; s.reason = 0; }
  if (leader) {
    *st_out = s;
    if (multi) *pp.seq = epoch0 + (unsigned long long)k; // one exchange per iteration that computed S p
  }
}

} // namespace gb

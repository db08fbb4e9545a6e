// direct_schur.cuh — direct solve of the reduced camera system S x = b_S on the GPU (GB_SOLVER_DIRECT_SCHUR).
//
// Replaces EigenSchurLDLTSolver::solve (solver/eigen_schur.hpp:72-108: S exported to a scalar upper CSC, copied to the
// HOST, Eigen::SimplicialLDLT factorisation and solve there, src/eigen_solver.cpp:10-29) and, by function,
// cudssSchurSolver (solver/cudss_schur.hpp:180-234).  S is SPD after damping, so a Cholesky factorisation gives the same
// step; it is done on the device on the dense form of S (9 Nc x 9 Nc, column-major, lower triangle), which is the
// practical regime of the reference's direct Schur solvers: a few hundred to ~2000 cameras.
//   k_dense_from_blocks   upper block-CSC values (explicit_schur.cuh: k_schur_build) -> dense lower triangle
//   k_cholesky            right-looking blocked Cholesky in ONE cooperative launch; per 32-column panel: diagonal block
//                         (one CTA, shared memory), panel solve (a thread per row), trailing update (32 x 32 tiles over
//                         the grid), three grid barriers.  A non-positive pivot sets the failure flag: the solve then
//                         reports solve_ok = false and the LM loop rejects the step (levenberg_marquardt.hpp:181-183),
//                         the reference's "Schur LDLT matrix decomposition failed" path.
//   k_cholesky_solve      L y = b_S, L^T x = y by one CTA, blocked by the same panels.
#pragma once
#include "kernels.cuh"

namespace gb {

constexpr int CH_NB = 32;       // panel width
constexpr int CH_THREADS = 256;

template <typename T>
__global__ void __launch_bounds__(256)
k_dense_from_blocks(int Nc, int nblocks, const int32_t *__restrict__ blk_row, const int32_t *__restrict__ blk_col,
                    const T *__restrict__ vals /*[nblocks][81] column-major*/, T *__restrict__ A /*[n][n] column-major*/) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (b >= nblocks) return;
  const int64_t n = 9 * (int64_t)Nc;
  const int i = blk_row[b], j = blk_col[b]; // i <= j: block (i, j) of the upper triangle = block (j, i)^T of the lower
  for (int e = lane; e < 81; e += 32) {
    const int r = e % 9, c = e / 9;
    A[(9 * (int64_t)j + c) + n * (9 * (int64_t)i + r)] = vals[(int64_t)b * 81 + e]; // element (9j+c, 9i+r) of the lower triangle
  }
}

template <typename T>
__global__ void __launch_bounds__(CH_THREADS)
k_cholesky(int n, T *A, int *fail) {
  namespace cg = cooperative_groups;
  cg::grid_group grid = cg::this_grid();
  __shared__ T L11[CH_NB][CH_NB + 1];
  __shared__ T Ta[CH_NB][CH_NB + 1], Tb[CH_NB][CH_NB + 1];
  const int tid = threadIdx.x, G = gridDim.x;
  const int64_t ld = n;
  for (int k0 = 0; k0 < n; k0 += CH_NB) {
    const int kb = min(CH_NB, n - k0);
    // ---- (a) diagonal block, CTA 0: unblocked Cholesky in shared memory ---------------------------------------------
    if (blockIdx.x == 0) {
      for (int e = tid; e < kb * kb; e += CH_THREADS) {
        const int r = e % kb, c = e / kb;
        L11[r][c] = r >= c ? __ldcg(A + (k0 + r) + ld * (k0 + c)) : T(0); // (updated by other CTAs: L2 loads)
      }
      __syncthreads();
      for (int c = 0; c < kb; c++) {
        const T d = L11[c][c];
        if (!(d > T(0))) { // not positive definite (or NaN): the reference's factorize() == false
          if (tid == 0) *fail = 1;
        }
        const T sq = sqrt(d);
        __syncthreads();
        if (tid < kb && tid >= c) L11[tid][c] = tid == c ? sq : L11[tid][c] / sq;
        __syncthreads();
        for (int e = tid; e < kb * kb; e += CH_THREADS) { // trailing part of the small block
          const int r = e % kb, cc = e / kb;
          if (cc > c && r >= cc) L11[r][cc] -= L11[r][c] * L11[cc][c];
        }
        __syncthreads();
      }
      for (int e = tid; e < kb * kb; e += CH_THREADS) {
        const int r = e % kb, c = e / kb;
        if (r >= c) A[(k0 + r) + ld * (k0 + c)] = L11[r][c];
      }
    }
    __threadfence();
    grid.sync();
    if (*(volatile int *)fail) return; // uniform: every CTA sees the flag after the barrier
    const int m0 = k0 + kb; // first row below the panel
    if (m0 >= n) break;
    // ---- (b) panel: L21 = A21 L11^-T, one thread per row ------------------------------------------------------------
    for (int e = tid; e < kb * kb; e += CH_THREADS) {
      const int r = e % kb, c = e / kb;
      L11[r][c] = r >= c ? __ldcg(A + (k0 + r) + ld * (k0 + c)) : T(0);
    }
    __syncthreads();
    for (int64_t i = m0 + (int64_t)blockIdx.x * CH_THREADS + tid; i < n; i += (int64_t)G * CH_THREADS) {
      T row[CH_NB];
#pragma unroll
      for (int c = 0; c < CH_NB; c++) row[c] = c < kb ? __ldcg(A + i + ld * (k0 + c)) : T(0);
#pragma unroll
      for (int c = 0; c < CH_NB; c++) {
        if (c < kb) {
          T v = row[c];
#pragma unroll
          for (int l = 0; l < CH_NB; l++)
            if (l < c) v -= row[l] * L11[c][l];
          row[c] = v / L11[c][c];
        }
      }
#pragma unroll
      for (int c = 0; c < CH_NB; c++)
        if (c < kb) A[i + ld * (k0 + c)] = row[c];
    }
    __threadfence();
    grid.sync();
    // ---- (c) trailing update A22 -= L21 L21^T, lower tiles of 32 x 32 dealt to the CTAs ------------------------------
    const int nt = (n - m0 + CH_NB - 1) / CH_NB;
    const int64_t ntiles = (int64_t)nt * (nt + 1) / 2;
    for (int64_t tix = blockIdx.x; tix < ntiles; tix += G) {
      // tile (ti, tj), tj <= ti, from the linear index of the lower triangle
      int ti = (int)((sqrt(8.0 * (double)tix + 1.0) - 1.0) * 0.5);
      while ((int64_t)ti * (ti + 1) / 2 > tix) ti--;
      while ((int64_t)(ti + 1) * (ti + 2) / 2 <= tix) ti++;
      const int tj = (int)(tix - (int64_t)ti * (ti + 1) / 2);
      const int r0 = m0 + ti * CH_NB, c0 = m0 + tj * CH_NB;
      __syncthreads();
      for (int e = tid; e < CH_NB * CH_NB; e += CH_THREADS) {
        const int r = e % CH_NB, l = e / CH_NB;
        Ta[r][l] = (r0 + r < n && l < kb) ? __ldcg(A + (r0 + r) + ld * (k0 + l)) : T(0);
        Tb[r][l] = (c0 + r < n && l < kb) ? __ldcg(A + (c0 + r) + ld * (k0 + l)) : T(0);
      }
      __syncthreads();
      for (int e = tid; e < CH_NB * CH_NB; e += CH_THREADS) {
        const int r = e % CH_NB, c = e / CH_NB;
        if (r0 + r < n && c0 + c < n && r0 + r >= c0 + c) {
          T acc = T(0);
#pragma unroll
          for (int l = 0; l < CH_NB; l++) acc += Ta[r][l] * Tb[c][l];
          T *dst = A + (r0 + r) + ld * (c0 + c);
          *dst = __ldcg(dst) - acc;
        }
      }
    }
    __threadfence();
    grid.sync();
  }
}

// x = (L L^T)^-1 b, one CTA.  y overwrites x in place.
template <typename T>
__global__ void __launch_bounds__(1024)
k_cholesky_solve(int n, const T *__restrict__ A, const T *__restrict__ b, T *__restrict__ x, const int *__restrict__ fail) {
  __shared__ T yb[CH_NB];
  const int tid = threadIdx.x;
  const int64_t ld = n;
  if (*fail) { // no factor: leave a zero step (the loop rejects it)
    for (int i = tid; i < n; i += blockDim.x) x[i] = T(0);
    return;
  }
  for (int i = tid; i < n; i += blockDim.x) x[i] = b[i];
  __syncthreads();
  // forward: L y = b
  for (int k0 = 0; k0 < n; k0 += CH_NB) {
    const int kb = min(CH_NB, n - k0);
    if (tid < 32) { // the 32 x 32 triangular system by one warp
      T v = tid < kb ? x[k0 + tid] : T(0);
      for (int c = 0; c < kb; c++) {
        const T piv = __shfl_sync(0xffffffffu, v, c) / A[(k0 + c) + ld * (k0 + c)];
        if (tid == c) v = piv;
        else if (tid > c && tid < kb) v -= A[(k0 + tid) + ld * (k0 + c)] * piv;
      }
      if (tid < kb) { x[k0 + tid] = v; yb[tid] = v; }
    }
    __syncthreads();
    for (int i = k0 + kb + tid; i < n; i += blockDim.x) {
      T acc = T(0);
      for (int l = 0; l < kb; l++) acc += A[i + ld * (k0 + l)] * yb[l];
      x[i] -= acc;
    }
    __syncthreads();
  }
  // backward: L^T x = y
  const int npan = (n + CH_NB - 1) / CH_NB;
  for (int pk = npan - 1; pk >= 0; pk--) {
    const int k0 = pk * CH_NB, kb = min(CH_NB, n - k0);
    if (tid < 32) {
      T v = tid < kb ? x[k0 + tid] : T(0);
      for (int c = kb - 1; c >= 0; c--) {
        const T piv = __shfl_sync(0xffffffffu, v, c) / A[(k0 + c) + ld * (k0 + c)];
        if (tid == c) v = piv;
        else if (tid < c) v -= A[(k0 + c) + ld * (k0 + tid)] * piv; // (L^T)[tid][c] = L[c][tid]
      }
      if (tid < kb) { x[k0 + tid] = v; yb[tid] = v; }
    }
    __syncthreads();
    for (int i = tid; i < k0; i += blockDim.x) {
      T acc = T(0);
      for (int l = 0; l < kb; l++) acc += A[(k0 + l) + ld * i] * yb[l]; // (L^T)[i][k0+l] = L[k0+l][i]
      x[i] -= acc;
    }
    __syncthreads();
  }
}

// what the host reads after a direct solve: no PCG iterations, stop reason 5 (solved) or 6 (factorisation failed)
template <typename T> __global__ void k_direct_state(PcgState<T> *st, const int *fail) {
  PcgState<T> s;
  s.rz = T(0); s.rz0 = T(0); s.alpha = T(0); s.beta = T(0); s.denom = T(0);
  s.iter = 0; s.done = 1; s.reason = *fail ? 6 : 5; s.pad = 0;
  *st = s;
}

} // namespace gb

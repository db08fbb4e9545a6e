// pcg_solve.cuh — the whole block-Jacobi PCG solve on the reduced camera system in ONE persistent cooperative kernel.
//
// Replaces PCGSchurSolver::solve (solver/pcg_schur.hpp:79-168) with execute_schur_vector_multiply (schur.hpp:347-393)
// and BlockJacobiSchurPreconditioner::apply (preconditioner/block_jacobi_schur.hpp:157-178) inside.  The reference
// runs ~8 launches, 3 stream synchronisations and 2 blocking scalar read-backs per PCG iteration; round 1 here ran 2
// launches (product + a cooperative update kernel with 2-3 grid barriers).  At 1/8 of Venice per GPU the product takes
// 40 us and those launches / barriers / exchanges took 33 us, so the solve is now one kernel:
//
//   grid = one CTA per SM (all co-resident, cooperative launch), 512 threads = the two workers of the product pipeline.
//   per PCG iteration k, TWO grid barriers:
//     phase P  every CTA forms p = beta p_old + z for the cameras it OWNS (and its share of sum p dterm p); then the
//              matrix-free product (B - E W E^T) D p over super-tiles handed out by an atomic work counter (an SM that
//              runs slow simply takes fewer), on the TMA pipeline of kernels.cuh whose ring keeps running across
//              super-tile boundaries.  The camera vector rows of a super-tile are built on the fly from p_old, z and
//              beta (no xs round trip).  Each super-tile also leaves  sum_rows (D p)_row . row  - its share of p.(S p) -
//              so the first dot product needs no pass over the reduced vectors.
//     barrier A
//              denom = sum of the super-tile dots (fixed order) + sum p dterm p ; alpha.   [multi-GPU: one scalar per
//              rank crosses NVLink here, overlapped with the row gather below]
//     phase U  owners: Ap_c = D_c sum(partial rows of c) [multi-GPU: pushed into every rank's receive slot with one
//              flag per CTA - CTA b only waits for CTA b of its peers - and added in rank order], x += alpha p,
//              r -= alpha Ap, z = M^-1 r, partial of r.z
//     barrier B
//              rz_new, rejection / convergence tests, beta           (pcg_schur.hpp:144-163)
//   All CTAs (and all ranks) carry the PCG scalars redundantly from bit-identical sums, so control flow is uniform
//   with no broadcast.  Every sum has a fixed order: runs are bit-reproducible whatever the work counter hands out.
#pragma once
#include "kernels.cuh"

namespace gb {

constexpr int SOLVE_THREADS = 2 * TILE, SOLVE_WARPS = SOLVE_THREADS / 32;
constexpr int SOLVE_STAMPS = 8; // per iteration: P start, before A, after A, exchange done, before B, after B

template <typename T, typename S> struct SolveSmem {
  using SM = SchurSmem2<T, S>;
  static constexpr int RED_OFF = SM::TOTAL;                       // T[32] reduction scratch
  static constexpr int WP_OFF = RED_OFF + 32 * (int)sizeof(double); // T[SOLVE_WARPS]
  static constexpr int CTL_OFF = WP_OFF + SOLVE_WARPS * (int)sizeof(double); // int[8]
  static constexpr int TOTAL = CTL_OFF + 64;
};

// p = beta p + z (ops::axpy_async(p, beta, p, z)): ONE definition, so that the owner of a camera and every CTA that
// rebuilds the camera's row of D p for the product round identically
template <typename T> __device__ __forceinline__ T pcg_direction(T beta, T p_old, T z) { return fma(beta, p_old, z); }

// per-warp values (lane 0) -> CTA total -> dst[blockIdx.x]
template <typename T> __device__ __forceinline__ void solve_publish(T wv, T *wpart, T *dst) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads(); // wpart may still be read from the previous use
  if (lane == 0) wpart[warp] = wv;
  __syncthreads();
  if (threadIdx.x == 0) {
    T tot = T(0);
#pragma unroll
    for (int w = 0; w < SOLVE_WARPS; w++) tot += wpart[w];
    *(volatile T *)(dst + blockIdx.x) = tot;
  }
}

// wait until every peer has published `epoch` in flag word f[q * stride]; a peer that never arrives sets the error flag
__device__ __forceinline__ void solve_wait_peers(const P2P &pp, const unsigned long long *f, int stride,
                                                 unsigned long long epoch) {
  if (threadIdx.x < pp.nranks && (int)threadIdx.x != pp.rank) {
    const unsigned long long *w = f + (size_t)threadIdx.x * stride;
    if (ld_acquire_sys(w) < epoch) {
      const unsigned long long t0 = global_timer_ns();
      while (ld_acquire_sys(w) < epoch) {
        if (global_timer_ns() - t0 > pp.timeout_ns) {
          *pp.error = 1;
          break;
        }
      }
    }
  }
  __syncthreads();
}

template <typename T, typename S>
__global__ void __launch_bounds__(SOLVE_THREADS, 1)
k_pcg_solve(DevStruct ds, const typename V2<S>::type *__restrict__ J, const T *__restrict__ W,
            const T *__restrict__ scale_c, const T *__restrict__ dterm, const T *__restrict__ Minv,
            const T *__restrict__ bS, T *x, T *xbak, T *r, T *z, T *pbuf /*[2][9 Nc]*/, T *part /*[nrows][9]*/,
            T *st_dot /*[nst]*/, T *cta_red /*[2][grid]*/, PcgState<T> *st_out, unsigned int *work, T tol, T ratio,
            int max_iter, P2P pp, int multi, unsigned long long *cta_flags /*own: [nranks][grid]*/,
            unsigned long long *timing /*[max_iter + 1][SOLVE_STAMPS] or null*/) {
  namespace cg = cooperative_groups;
  using SM = SchurSmem2<T, S>;
  using SS = SolveSmem<T, S>;
  using S2 = typename V2<S>::type;
  cg::grid_group grid = cg::this_grid();
  extern __shared__ __align__(128) unsigned char smem[];
  const int worker = threadIdx.x >> 8, t = threadIdx.x & (TILE - 1);
  T *xl = reinterpret_cast<T *>(smem + SM::XL_OFF);
  T *acc_all = reinterpret_cast<T *>(smem + SM::ACC_OFF);
  T *acc = acc_all + worker * SLOT_CAP * 9;
  T *sv = reinterpret_cast<T *>(smem + SM::STG_OFF + worker * SM::STG_BYTES);
  T *sw = reinterpret_cast<T *>(smem + SM::STG_OFF + worker * SM::STG_BYTES + SM::SV_BYTES);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + SM::BAR_OFF);
  T *red = reinterpret_cast<T *>(smem + SS::RED_OFF);
  T *wpart = reinterpret_cast<T *>(smem + SS::WP_OFF);
  volatile int *ctl = reinterpret_cast<volatile int *>(smem + SS::CTL_OFF);
  const uint64_t pol = l2_policy_evict_first();

  const int G = gridDim.x, Nc = ds.Nc, dimc = 9 * Nc, nst = ds.nst;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool leader = blockIdx.x == 0 && threadIdx.x == 0;
  // camera ownership: contiguous blocks of cameras per CTA, one warp per camera at a time; lanes 0..8 own the nine
  // entries, lanes (sub, k) = (lane / 9, lane % 9), lane < 27, share the row gather
  const int cpc = (Nc + G - 1) / G;
  const int c_begin = min(Nc, (int)blockIdx.x * cpc), c_end = min(Nc, c_begin + cpc);
  const int k9 = lane % 9, sub = lane / 9;
  const bool own = lane < 9;
  // epoch of the last exchange before this solve: read by everybody before anybody can advance it (end of kernel)
  const unsigned long long epoch0 = multi ? p2p_current_epoch(pp) : 0ull;
  const unsigned int grabs_per_iter = (unsigned int)(nst + G); // every CTA ends an iteration with one failing grab

  if (threadIdx.x == 0) {
    for (int s = 0; s < SM::NMETA; s++) mbar_init(&bars[s], 1);
    mbar_fence_init();
    fence_proxy_async();
  }
  if (leader) *work = 0u;

  // ---- start: x = 0, r = b_S, z = M^-1 r, p_old = 0, rz = r.z   (pcg_schur.hpp:90-106) ----------------------------
  {
    T rz_w = T(0);
    for (int c = c_begin + warp; c < c_end; c += SOLVE_WARPS) {
      const int i = c * 9 + k9;
      T rn = T(0);
      if (own) {
        rn = bS[i];
        x[i] = T(0);
        r[i] = rn;
        pbuf[i] = T(0);
      }
      const T *m = Minv + (int64_t)c * 81;
      T a = T(0);
#pragma unroll
      for (int j = 0; j < 9; j++) {
        const T rj = __shfl_sync(0xffffffffu, rn, j);
        if (own) a += m[k9 + 9 * j] * rj;
      }
      if (own) z[i] = a;
      rz_w += sum9<T>(own ? rn * a : T(0));
    }
    solve_publish<T>(rz_w, wpart, cta_red + G);
  }
  __threadfence();
  grid.sync();
  PcgState<T> s;
  s.rz = grid_total<T>(cta_red + G, G, red);
  s.rz0 = (T)INFINITY; s.alpha = T(0); s.beta = T(0); s.denom = T(0);
  s.iter = 0; s.done = 0; s.reason = 0; s.pad = 0;
  T beta = T(0);
  int ring = 0;        // ring index of the next tile of this CTA's tile sequence
  int my_issued = -1;  // (thread 0 of each worker) highest ring index of the worker's parity whose copies are issued
  int k = 0;

  for (; k < max_iter; k++) {
    if (s.rz == T(0)) { s.done = 1; s.reason = 3; break; } // pcg_schur.hpp:109-111
    if (timing && leader) timing[k * SOLVE_STAMPS + 0] = global_timer_ns();
    const T *p_old = pbuf + (size_t)(k & 1) * dimc;
    T *p_new = pbuf + (size_t)((k + 1) & 1) * dimc;
    // ---- phase P: p = beta p + z for the owned cameras (ops::axpy_async(p, beta, p, z)), sum p dterm p --------------
    {
      T pdp_w = T(0);
      for (int c = c_begin + warp; c < c_end; c += SOLVE_WARPS) {
        T v = T(0);
        if (own) {
          const int i = c * 9 + k9;
          const T pn = pcg_direction<T>(beta, p_old[i], z[i]);
          p_new[i] = pn;
          v = pn * (dterm[i] * pn);
        }
        pdp_w += sum9<T>(v);
      }
      solve_publish<T>(pdp_w, wpart, cta_red);
    }
    // ---- the product over the super-tiles the work counter hands out -----------------------------------------------
    {
      const unsigned int base = (unsigned int)k * grabs_per_iter;
      if (threadIdx.x == 0) ctl[0] = (int)(atomicAdd(work, 1u) - base);
      __syncthreads();
      int cur = ctl[0];
      int ib = ring;
      while (cur < nst) {
        if (threadIdx.x == 0) ctl[1] = (int)(atomicAdd(work, 1u) - base); // the next item, known a super-tile ahead
        const int tile_base = ds.st_tile[cur], ie = ib + (ds.st_tile[cur + 1] - tile_base);
        const int row0 = ds.st_row[cur], nslots = ds.st_row[cur + 1] - row0;
        for (int i = threadIdx.x; i < nslots * 9; i += SOLVE_THREADS) {
          const int sl = i / 9, kk = i - 9 * sl;
          const int j = ds.row_cam[row0 + sl] * 9 + kk;
          // (D p)_c with the owner's arithmetic; p_old and z were written by other CTAs during this kernel: L2 loads
          xl[i] = scale_c[j] * pcg_direction<T>(beta, __ldcg(p_old + j), __ldcg(z + j));
        }
        for (int i = t; i < nslots * 9; i += TILE) acc[i] = T(0);
        __syncthreads();
        const int nxt = ctl[1];
        int nb_tile = 0, nb_nt = 0;
        if (nxt < nst) { nb_tile = ds.st_tile[nxt]; nb_nt = ds.st_tile[nxt + 1] - nb_tile; }
        for (int i = ib + ((ib ^ worker) & 1); i < ie; i += 2) { // worker w takes ring indices of parity w
          if (t == 0 && i > my_issued) { // not prefetched (first tiles of an iteration, one-tile super-tiles)
            const int tile = tile_base + (i - ib);
            const int p0 = ds.tmeta[tile].p0, np = ds.tmeta[tile].np;
            fence_proxy_async();
            product_issue<T, S>(smem, bars, ds, J, W, tile, i, p0, np, pol);
            my_issued = i;
          }
          mbar_wait(&bars[i & 3], (uint32_t)((i >> 2) & 1));
          const S2 *Js = reinterpret_cast<const S2 *>(smem + (i & 1) * SM::J_BYTES);
          const unsigned char *rec = smem + SM::META_OFF + (i & 3) * SM::META_BYTES;
          const T *Ws = reinterpret_cast<const T *>(rec + REC_BYTES);
          product_tile<T, S, false>(worker, t, Js, rec, Ws, xl, acc, sv, sw, (T *)nullptr, [&](int next_p0, int next_np) {
            const int j = i + 2;
            if (j <= my_issued) return;
            if (j < ie) {
              fence_proxy_async();
              product_issue<T, S>(smem, bars, ds, J, W, tile_base + (j - ib), j, next_p0, next_np, pol);
              my_issued = j;
            } else if (j - ie < nb_nt) { // first tiles of the NEXT super-tile: the ring runs across the boundary
              const int tile = nb_tile + (j - ie);
              const int p0 = ds.tmeta[tile].p0, np = ds.tmeta[tile].np;
              fence_proxy_async();
              product_issue<T, S>(smem, bars, ds, J, W, tile, j, p0, np, pol);
              my_issued = j;
            }
          });
        }
        __syncthreads();
        // rows of this super-tile (worker 0 + worker 1, fixed order) and its share of p . (S p)
        T d = T(0);
        for (int i = threadIdx.x; i < nslots * 9; i += SOLVE_THREADS) {
          const int sl = i / 9, kk = i - 9 * sl;
          const T v = acc_all[i] + acc_all[SLOT_CAP * 9 + i];
          part[(int64_t)ds.row_out[row0 + sl] * 9 + kk] = v;
          d += xl[i] * v;
        }
        d = block_sum<T>(d, red);
        if (threadIdx.x == 0) st_dot[cur] = d;
        __syncthreads(); // xl / acc / ctl are rewritten by the next super-tile
        ib = ie;
        cur = nxt;
      }
      ring = ib;
    }
    if (timing && leader) timing[k * SOLVE_STAMPS + 1] = global_timer_ns();
    __threadfence();
    grid.sync(); // ---- barrier A: partial rows, super-tile dots and p are visible -------------------------------------
    if (timing && leader) timing[k * SOLVE_STAMPS + 2] = global_timer_ns();
    const unsigned long long epoch = epoch0 + (unsigned long long)k + 1ull;
    T dot = grid_total<T>(st_dot, nst, red);
    const T pdp = grid_total<T>(cta_red, G, red);
    if (multi && blockIdx.x == 0) {
      // this rank's share of p.(S p) goes to every peer (one scalar after the vector area of the slot), then the flag
      if (threadIdx.x < pp.nranks && (int)threadIdx.x != pp.rank)
        *reinterpret_cast<T *>(reinterpret_cast<unsigned char *>(p2p_slot<T>(pp, threadIdx.x, pp.rank, epoch)) +
                               (size_t)54 * Nc * sizeof(T)) = dot;
      __threadfence_system();
      __syncthreads();
      if (threadIdx.x < pp.nranks && (int)threadIdx.x != pp.rank) st_relaxed_sys(pp.flags[threadIdx.x] + pp.rank, epoch);
    }

    // ---- phase U: Ap_raw = D_c * (sum of the camera's partial rows) --------------------------------------------------
    auto gather_raw = [&](int c) -> T { // valid in lanes 0..8
      const int b = ds.cam_row_ptr[c], n = ds.cam_row_ptr[c + 1] - b;
      T a0 = T(0), a1 = T(0), a2 = T(0), a3 = T(0), a4 = T(0), a5 = T(0), a6 = T(0), a7 = T(0);
      if (sub < 3) {
        const T *base = part + (int64_t)b * 9 + k9; // written by other CTAs during this kernel: L2 loads
        int row = sub;
        for (; row + 21 < n; row += 24) { // eight independent loads in flight per lane
          a0 += __ldcg(base + row * 9);
          a1 += __ldcg(base + (row + 3) * 9);
          a2 += __ldcg(base + (row + 6) * 9);
          a3 += __ldcg(base + (row + 9) * 9);
          a4 += __ldcg(base + (row + 12) * 9);
          a5 += __ldcg(base + (row + 15) * 9);
          a6 += __ldcg(base + (row + 18) * 9);
          a7 += __ldcg(base + (row + 21) * 9);
        }
        for (; row + 3 < n; row += 6) {
          a0 += __ldcg(base + row * 9);
          a1 += __ldcg(base + (row + 3) * 9);
        }
        for (; row < n; row += 3) a0 += __ldcg(base + row * 9);
      }
      const T v = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
      const T v1 = __shfl_sync(0xffffffffu, v, k9 + 9), v2 = __shfl_sync(0xffffffffu, v, k9 + 18);
      return own ? scale_c[c * 9 + k9] * ((v + v1) + v2) : T(0);
    };
    if (multi) {
      // this rank's sums go straight into every rank's receive slot; one flag per CTA: CTA b of a peer owns the same
      // cameras and waits for nothing else
      for (int c = c_begin + warp; c < c_end; c += SOLVE_WARPS) {
        const T raw = gather_raw(c);
        if (own)
          for (int q = 0; q < pp.nranks; q++) p2p_slot<T>(pp, q, pp.rank, epoch)[c * 9 + k9] = raw;
      }
      __threadfence_system();
      __syncthreads();
      if (threadIdx.x < pp.nranks && (int)threadIdx.x != pp.rank)
        st_relaxed_sys(reinterpret_cast<unsigned long long *>(pp.recv[threadIdx.x] + pp.cta_flag_off) +
                           (size_t)pp.rank * G + blockIdx.x, epoch);
      solve_wait_peers(pp, pp.flags[pp.rank], 1, epoch);
      T tot = T(0);
      for (int q = 0; q < pp.nranks; q++) // rank order; this rank's own share straight from the register
        tot += q == pp.rank ? dot
                            : __ldcg(reinterpret_cast<const T *>(reinterpret_cast<const unsigned char *>(
                                                                     p2p_slot<T>(pp, pp.rank, q, epoch)) +
                                                                 (size_t)54 * Nc * sizeof(T)));
      dot = tot;
      solve_wait_peers(pp, cta_flags + blockIdx.x, G, epoch);
    }
    if (timing && leader) timing[k * SOLVE_STAMPS + 3] = global_timer_ns();
    const T denom = dot + pdp;
    if (denom == T(0) || isnan(denom)) { s.done = 1; s.reason = 4; s.denom = denom; break; } // pcg_schur.hpp:120-122
    const T alpha = s.rz / denom;

    // x += alpha p ; r -= alpha Ap ; z = M^-1 r ; r.z
    T rz_w = T(0);
    for (int c = c_begin + warp; c < c_end; c += SOLVE_WARPS) {
      const int i = c * 9 + k9;
      const T raw0 = multi ? T(0) : gather_raw(c);
      T rn = T(0);
      if (own) {
        const T raw = multi ? p2p_sum<T>(pp, epoch, i) : raw0;
        const T pn = p_new[i];
        const T ap = raw + dterm[i] * pn;
        const T xo = x[i];
        xbak[i] = xo;
        x[i] = alpha * pn + xo;   // ops::axpy_async(x, alpha, p, x)
        rn = -alpha * ap + r[i];  // ops::axpy_async(r, -alpha, Ap, r)
        r[i] = rn;
      }
      const T *m = Minv + (int64_t)c * 81;
      T a = T(0);
#pragma unroll
      for (int j = 0; j < 9; j++) {
        const T rj = __shfl_sync(0xffffffffu, rn, j);
        if (own) a += m[k9 + 9 * j] * rj;
      }
      if (own) z[i] = a;
      rz_w += sum9<T>(own ? rn * a : T(0));
    }
    solve_publish<T>(rz_w, wpart, cta_red + G);
    if (timing && leader) timing[k * SOLVE_STAMPS + 4] = global_timer_ns();
    __threadfence();
    grid.sync(); // ---- barrier B: z and the r.z partials are visible --------------------------------------------------
    if (timing && leader) timing[k * SOLVE_STAMPS + 5] = global_timer_ns();
    const T rzn = grid_total<T>(cta_red + G, G, red);
    s.iter += 1;
    s.alpha = alpha;
    s.denom = denom;
    if (fabs(rzn) > ratio * s.rz0 || isnan(rzn)) { // pcg_schur.hpp:144-148: restore x, stop
      for (int c = c_begin + warp; c < c_end; c += SOLVE_WARPS)
        if (own) x[c * 9 + k9] = xbak[c * 9 + k9];
      s.done = 1; s.reason = 2; s.rz = rzn;
      k++;
      break;
    }
    s.rz0 = fmin(s.rz0, fabs(rzn));
    beta = rzn / s.rz;
    s.beta = beta;
    s.rz = rzn;
    if (fabs(rzn) < tol) { s.done = 1; s.reason = 1; k++; break; } // :160-162 (the p update before it is not needed)
  }
  if (!s.done) { s.done = 1; s.reason = 0; }
  if (leader) {
    *st_out = s;
    // exchanges consumed: one per iteration that reached barrier A (k counts them on every exit path)
    if (multi) *pp.seq = epoch0 + (unsigned long long)(s.reason == 4 ? k + 1 : k);
  }
}

} // namespace gb

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
//              owners: Ap_c = D_c sum(partial rows of c)  [multi-GPU: stored straight into every rank's LL slot - data and
//              epoch in one 8-byte word, no fence, no flag (p2p.cuh); CTA b of a peer owns the same cameras and polls
//              exactly those words; the sums are added in rank order]
//              denom = sum of the super-tile dots (fixed order) + sum p dterm p ; alpha   [multi-GPU: plus one scalar
//              per rank, sent the same way while the vector is in flight]
//     phase U  x += alpha p, r -= alpha Ap, z = M^-1 r, partial of r.z
//     barrier B
//              rz_new, rejection / convergence tests, beta           (pcg_schur.hpp:144-163)
//   With long tracks (points cut into fragment tiles, structure.hpp) a phase H precedes the product: one warp per such point
//   forms t_p over all its observations, then one more grid barrier.  Solves of <= 16 iterations also keep every point sum
//   t_p(p_k) and step length alpha_k: the back-substitution then needs no pass over the Jacobians (k_backsubst_points).
//   All CTAs (and all ranks) carry the PCG scalars redundantly from bit-identical sums, so control flow is uniform
//   with no broadcast.  Every sum has a fixed order: runs are bit-reproducible whatever the work counter hands out.
#pragma once
#include "kernels.cuh"

namespace gb {

// workers per CTA of the solve kernel: three for (float, float) - 80 registers without spills, 178 KB of shared memory -
// two otherwise (FP64 staging and Jacobian slots fill the 227 KB with two)
template <typename T, typename S> struct SolveWorkers { static constexpr int value = (sizeof(T) == 4 && sizeof(S) == 4) ? 3 : 2; };
constexpr int SOLVE_STAMPS = 8; // per iteration: P start, before A, after A, exchange done, before B, after B, [6] pushed

template <typename T, typename S> struct SolveSmem {
  static constexpr int NW = SolveWorkers<T, S>::value;
  static constexpr int THREADS = NW * TILE, WARPS = THREADS / 32;
  using SM = SchurSmem2<T, S, NW>;
  static constexpr int RED_OFF = SM::TOTAL;                       // T[64] reduction scratch
  static constexpr int WP_OFF = RED_OFF + 64 * (int)sizeof(double); // T[WARPS]
  static constexpr int CTL_OFF = WP_OFF + 32 * (int)sizeof(double); // int[8]
  static constexpr int NREC = 3;                                  // ring of per-super-tile records (current, next, next-next)
  static constexpr int REC_OFF = CTL_OFF + 64;
  static constexpr int TOTAL = REC_OFF + NREC * STREC_BYTES;
  static_assert(REC_OFF % 16 == 0 && STREC_BYTES % 16 == 0, "TMA destinations must stay 16-byte aligned");
};

// p = beta p + z (ops::axpy_async(p, beta, p, z)): ONE definition, so that the owner of a camera and every CTA that
// rebuilds the camera's row of D p for the product round identically
template <typename T> __device__ __forceinline__ T pcg_direction(T beta, T p_old, T z) { return fma(beta, p_old, z); }

// per-warp values (lane 0) -> CTA total -> dst[blockIdx.x]
template <typename T, int SOLVE_WARPS> __device__ __forceinline__ void solve_publish(T wv, T *wpart, T *dst) {
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

template <typename T, typename S>
__global__ void __launch_bounds__((SolveSmem<T, S>::THREADS), 1)
k_pcg_solve(DevStruct ds, const typename V2<S>::type *__restrict__ J, const T *__restrict__ W,
            const T *__restrict__ scale_c, const T *__restrict__ dterm, const T *__restrict__ Minv,
            const T *__restrict__ bS, T *x, T *xbak, T *r, T *z, T *pbuf /*[2][9 Nc]*/, T *part /*[nrows][9]*/,
            T *st_dot /*[nst]*/, T *cta_red /*[2][grid]*/, PcgState<T> *st_out, unsigned int *work, T tol, T ratio,
            int max_iter, P2P pp, int multi, unsigned long long *timing /*[max_iter + 1][SOLVE_STAMPS] or null*/) {
  namespace cg = cooperative_groups;
  using SS = SolveSmem<T, S>;
  using SM = typename SS::SM;
  constexpr int NW = SS::NW, SOLVE_THREADS = SS::THREADS, SOLVE_WARPS = SS::WARPS, NMETA = SM::NMETA;
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

  unsigned char *recbuf = smem + SS::REC_OFF;
  uint64_t *rbar = bars + SM::NMETA; // one mbarrier per record buffer
  if (threadIdx.x == 0) {
    for (int s = 0; s < SM::NMETA + SS::NREC; s++) mbar_init(&bars[s], 1);
    mbar_fence_init();
    fence_proxy_async();
  }
  if (leader) {
    *work = 0u;
    for (int i = 0; i < ds.tk_cap; i++) reinterpret_cast<T *>(ds.tk_alpha)[i] = T(0); // step lengths of the iterates applied to x
  }
  // Work items (super-tiles) come from an atomic counter that is never reset during the solve: every CTA ends an
  // iteration with exactly ONE failing grab, so iteration k hands out the values k * (nst + G) ... .  Item number n of this
  // CTA (over the whole solve) has its record in buffer n % 3, on that buffer's mbarrier with phase parity (n / 3) & 1.
  int n_cur = 0; // sequence number of the current item (identical in all threads)
  // thread 0: grab the first two items of iteration kk and start fetching their records; ctl[0] = cur, ctl[1] = next
  auto grab_first = [&](int kk) {
    const unsigned int base = (unsigned int)kk * grabs_per_iter;
    const int cur = (int)(atomicAdd(work, 1u) - base);
    const int nxt = cur < nst ? (int)(atomicAdd(work, 1u) - base) : cur;
    fence_proxy_async();
    if (cur < nst) {
      mbar_expect_tx(&rbar[n_cur % SS::NREC], STREC_BYTES);
      bulk_g2s(recbuf + (n_cur % SS::NREC) * STREC_BYTES, ds.strec + (int64_t)cur * STREC_BYTES, STREC_BYTES, &rbar[n_cur % SS::NREC]);
    }
    if (nxt < nst) {
      mbar_expect_tx(&rbar[(n_cur + 1) % SS::NREC], STREC_BYTES);
      bulk_g2s(recbuf + ((n_cur + 1) % SS::NREC) * STREC_BYTES, ds.strec + (int64_t)nxt * STREC_BYTES, STREC_BYTES,
               &rbar[(n_cur + 1) % SS::NREC]);
    }
    ctl[0] = cur;
    ctl[1] = nxt;
  };
  auto rec_wait = [&](int n) { mbar_wait(&rbar[n % SS::NREC], (uint32_t)((n / SS::NREC) & 1)); };
  auto rec_ptr = [&](int n) { return reinterpret_cast<const int32_t *>(recbuf + (n % SS::NREC) * STREC_BYTES); };

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
    solve_publish<T, SOLVE_WARPS>(rz_w, wpart, cta_red + G);
  }
  __threadfence();
  grid.sync();
  PcgState<T> s;
  s.rz = grid_total<T>(cta_red + G, G, red);
  s.rz0 = (T)INFINITY; s.alpha = T(0); s.beta = T(0); s.denom = T(0);
  s.iter = 0; s.done = 0; s.reason = 0; s.pad = 0;
  T beta = T(0);
  if (threadIdx.x == 0) grab_first(0); // (later iterations: grabbed before barrier B, while other CTAs still update)
  bool pregrabbed = true;              // items are grabbed and their records in flight, not yet consumed
  int ring = 0;        // ring index of the next tile of this CTA's tile sequence
  int my_issued = -1;  // (thread 0 of each worker) highest ring index of the worker's parity whose copies are issued
  int k = 0;

  for (; k < max_iter; k++) {
    if (s.rz == T(0)) { s.done = 1; s.reason = 3; break; } // pcg_schur.hpp:109-111
    if (timing && leader) timing[k * SOLVE_STAMPS + 0] = global_timer_ns();
    const T *p_old = pbuf + (size_t)(k & 1) * dimc;
    T *p_new = pbuf + (size_t)((k + 1) & 1) * dimc;
    const int keep_k = k < ds.tk_cap ? k : -1; // this iteration's point sums are kept for k_backsubst_points
    T *tk_cur = keep_k >= 0 ? reinterpret_cast<T *>(ds.tk) + (size_t)k * 3 * (size_t)ds.Np : nullptr;
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
      solve_publish<T, SOLVE_WARPS>(pdp_w, wpart, cta_red);
    }
    // ---- phase H (only with long tracks, structure.hpp): t_p = sum_o Jp^T Jc (D p)_c over ALL observations of every point
    // that is cut into fragment tiles (one warp per point), before any of its fragments is multiplied; one more grid barrier --
    if (ds.nheavy > 0) {
      for (int hp = warp * G + (int)blockIdx.x; hp < ds.nheavy; hp += G * SOLVE_WARPS) // one warp per point, spread over the SMs
        heavy_point_dot<T, S>(ds, J, hp, [&](int c, int j) {
          const int i = c * 9 + j;
          return scale_c[i] * pcg_direction<T>(beta, __ldcg(p_old + i), __ldcg(z + i));
        });
      __threadfence();
      grid.sync();
    }
    // ---- the product over the super-tiles the work counter hands out -----------------------------------------------
    // A super-tile boundary costs two CTA barriers and one pass: the rows and the dot of the finished super-tile are
    // written in the same pass that builds the camera vector rows of the next one, whose record (camera list, row
    // positions, tile range) was fetched by TMA a super-tile ahead.  (First version: separate epilogue and prologue with
    // dependent global loads st_row -> row_cam -> p, z and five barriers: 3.6 us per boundary, measured by sweeping the
    // super-tile size, gpurun_out/r2h_sweep4.log.)
    {
      const unsigned int base = (unsigned int)k * grabs_per_iter;
      // xl row i of camera cam[i / 9]: (D p)_c with the owner's arithmetic; p_old and z were written by other CTAs
      // during this kernel: L2 loads
      auto build_xl = [&](const int32_t *cam, int i) {
        const int sl = i / 9, kk = i - 9 * sl;
        const int j = cam[sl] * 9 + kk;
        xl[i] = scale_c[j] * pcg_direction<T>(beta, __ldcg(p_old + j), __ldcg(z + j));
      };
      __syncthreads(); // ctl[0], ctl[1] of grab_first
      pregrabbed = false;
      int cur = ctl[0], nxt = ctl[1];
      int ib = ring;
      if (cur < nst) {
        rec_wait(n_cur);
        const int32_t *R = rec_ptr(n_cur);
        const int n9 = R[3] * 9;
        for (int i = threadIdx.x; i < n9; i += SOLVE_THREADS) {
          build_xl(R + STREC_CAM / 4, i);
#pragma unroll
          for (int w = 0; w < NW; w++) acc_all[w * SLOT_CAP * 9 + i] = T(0);
        }
        __syncthreads();
      }
      while (cur < nst) {
        const int32_t *R = rec_ptr(n_cur);
        const int tile_base = R[0], ie = ib + R[1], n9 = R[3] * 9;
        for (int i = ib + (worker - ib % NW + NW) % NW; i < ie; i += NW) { // worker w takes the ring indices i with i % NW == w
          if (t == 0 && i > my_issued) { // not prefetched (first tiles of an iteration, one-tile super-tiles)
            const int tile = tile_base + (i - ib);
            const int p0 = ds.tmeta[tile].p0, np = ds.tmeta[tile].np;
            fence_proxy_async();
            product_issue<T, S, NW>(smem, bars, ds, J, W, tile, i, p0, np, pol);
            my_issued = i;
          }
          mbar_wait(&bars[i % NMETA], (uint32_t)((i / NMETA) & 1));
          const S2 *Js = reinterpret_cast<const S2 *>(smem + (i % NW) * SM::J_BYTES);
          const unsigned char *rec = smem + SM::META_OFF + (i % NMETA) * SM::META_BYTES;
          const T *Ws = reinterpret_cast<const T *>(rec + REC_BYTES);
          product_tile<T, S, false, NW>(worker, t, Js, rec, Ws, xl, acc, sv, sw, (T *)nullptr, ds, tk_cur, [&](int next_p0, int next_np) {
            const int j = i + NW;
            if (j <= my_issued) return;
            if (j < ie) {
              fence_proxy_async();
              product_issue<T, S, NW>(smem, bars, ds, J, W, tile_base + (j - ib), j, next_p0, next_np, pol);
              my_issued = j;
            } else if (nxt < nst) { // first tiles of the NEXT super-tile: the ring runs across the boundary
              rec_wait(n_cur + 1);
              const int32_t *RN = rec_ptr(n_cur + 1);
              if (j - ie < RN[1]) {
                const int tile = RN[0] + (j - ie);
                const int p0 = ds.tmeta[tile].p0, np = ds.tmeta[tile].np;
                fence_proxy_async();
                product_issue<T, S, NW>(smem, bars, ds, J, W, tile, j, p0, np, pol);
                my_issued = j;
              }
            }
          });
        }
        __syncthreads(); // both workers' accumulator rows are complete
        // the item after next, grabbed a super-tile ahead (its record lands while the next super-tile is processed)
        if (threadIdx.x == 0) {
          int nn = nxt;
          if (nxt < nst) {
            nn = (int)(atomicAdd(work, 1u) - base);
            if (nn < nst) {
              fence_proxy_async();
              mbar_expect_tx(&rbar[(n_cur + 2) % SS::NREC], STREC_BYTES);
              bulk_g2s(recbuf + ((n_cur + 2) % SS::NREC) * STREC_BYTES, ds.strec + (int64_t)nn * STREC_BYTES, STREC_BYTES,
                       &rbar[(n_cur + 2) % SS::NREC]);
            }
          }
          ctl[2] = nn;
        }
        // one pass: rows of this super-tile (worker 0 + worker 1, fixed order), its share of p . (S p), and the camera
        // vector rows + cleared accumulators of the next super-tile
        const int32_t *RN = nullptr;
        int m9 = 0;
        if (nxt < nst) {
          rec_wait(n_cur + 1);
          RN = rec_ptr(n_cur + 1);
          m9 = RN[3] * 9;
        }
        T d = T(0);
        const int32_t *out = R + STREC_OUT / 4;
        for (int i = threadIdx.x; i < max(n9, m9); i += SOLVE_THREADS) {
          if (i < n9) {
            const int sl = i / 9, kk = i - 9 * sl;
            T v = acc_all[i] + acc_all[SLOT_CAP * 9 + i]; // the workers' rows in worker order
#pragma unroll
            for (int w = 2; w < NW; w++) v += acc_all[w * SLOT_CAP * 9 + i];
            part[(int64_t)out[sl] * 9 + kk] = v;
            d += xl[i] * v;
          }
          if (i < m9) {
            build_xl(RN + STREC_CAM / 4, i);
#pragma unroll
            for (int w = 0; w < NW; w++) acc_all[w * SLOT_CAP * 9 + i] = T(0);
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) d += __shfl_down_sync(0xffffffffu, d, o);
        if (lane == 0) wpart[warp] = d;
        __syncthreads(); // next super-tile's rows are ready; the warp partials of the dot and ctl[2] are visible
        if (threadIdx.x == 0) {
          T tot = T(0);
#pragma unroll
          for (int w = 0; w < SOLVE_WARPS; w++) tot += wpart[w];
          st_dot[cur] = tot;
        }
        ib = ie;
        cur = nxt;
        nxt = ctl[2];
        n_cur++;
      }
      ring = ib;
    }
    if (timing && leader) timing[k * SOLVE_STAMPS + 1] = global_timer_ns();
    __threadfence();
    grid.sync(); // ---- barrier A: partial rows, super-tile dots and p are visible -------------------------------------
    if (timing && leader) timing[k * SOLVE_STAMPS + 2] = global_timer_ns();
    const unsigned long long epoch = epoch0 + (unsigned long long)k + 1ull;
    const unsigned int e32 = (unsigned int)epoch;
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
      // this rank's sums go straight into every rank's LL slot (p2p.cuh): data and epoch in one 8-byte word, no fence, no
      // flag; CTA b of every peer owns the same cameras and polls exactly these words
      for (int c = c_begin + warp; c < c_end; c += SOLVE_WARPS) {
        const T raw = gather_raw(c);
        if (own)
          for (int q = 0; q < pp.nranks; q++) ll_store(ll_slot(pp, q, pp.rank, epoch), (long long)c * 9 + k9, raw, e32);
      }
    }
    T dot, pdp;
    grid_total2<T>(st_dot, nst, cta_red, G, red, dot, pdp);
    if (multi) {
      // p.(S p) needs one scalar per rank: CTA 0 sends this rank's (the word after the vector), everybody reads all of them
      if (blockIdx.x == 0 && threadIdx.x < pp.nranks) // (to this rank's own slot too: ll_sum reads all nranks slots)
        ll_store(ll_slot(pp, threadIdx.x, pp.rank, epoch), (long long)9 * Nc, dot, e32);
      if (timing && leader) timing[k * SOLVE_STAMPS + 6] = global_timer_ns();
      dot = ll_sum<T>(pp, epoch, (long long)9 * Nc, e32); // rank order: bit-identical on every rank
    }
    if (timing && leader) timing[k * SOLVE_STAMPS + 3] = global_timer_ns();
    const T denom = dot + pdp;
    if (denom == T(0) || isnan(denom)) { s.done = 1; s.reason = 4; s.denom = denom; break; } // pcg_schur.hpp:120-122
    const T alpha = s.rz / denom;
    if (leader && keep_k >= 0) reinterpret_cast<T *>(ds.tk_alpha)[k] = alpha;

    // x += alpha p ; r -= alpha Ap ; z = M^-1 r ; r.z
    T rz_w = T(0);
    for (int c = c_begin + warp; c < c_end; c += SOLVE_WARPS) {
      const int i = c * 9 + k9;
      const T raw0 = multi ? T(0) : gather_raw(c);
      T rn = T(0);
      if (own) {
        // multi-GPU: the ranks' sums in rank order (this rank's own went through its own slot too)
        const T raw = multi ? ll_sum<T>(pp, epoch, i, e32) : raw0;
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
    // this CTA's share of the update is done: take the first work items of the next iteration now, so that their
    // records are in shared memory when barrier B opens
    if (k + 1 < max_iter) {
      if (threadIdx.x == 0) grab_first(k + 1);
      pregrabbed = true;
    }
    solve_publish<T, SOLVE_WARPS>(rz_w, wpart, cta_red + G);
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
      if (leader && keep_k >= 0) reinterpret_cast<T *>(ds.tk_alpha)[k] = T(0); // the rejected iterate is not part of x
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
  if (pregrabbed) { // records of items that will never be processed are still in flight: let them land before the CTA exits
    __syncthreads();
    if (ctl[0] < nst) rec_wait(n_cur);
    if (ctl[1] < nst) rec_wait(n_cur + 1);
  }
  if (leader) {
    *st_out = s;
    // exchanges consumed: one per iteration that reached barrier A (k counts them on every exit path)
    if (multi) *pp.seq = epoch0 + (unsigned long long)(s.reason == 4 ? k + 1 : k);
  }
}

} // namespace gb

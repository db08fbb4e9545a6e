// api.cu — context, problem state, host-side LM driver and the C ABI (include/graphite_b200.h).
//
// Host control flow mirrors optimizer::levenberg_marquardt (include/graphite/optimizer/levenberg_marquardt.hpp:109-242)
// and PCGSchurSolver (include/graphite/solver/pcg_schur.hpp:49-168); all arithmetic runs in the kernels of
// kernels.cuh.  There is no CPU fallback: every entry point needs a live CUDA context.
#include "context.hpp"

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <dlfcn.h>
#include <limits>
#include <string>
#include <vector>

#include "kernels.cuh"
#include "pcg_solve.cuh"
#include "explicit_schur.cuh"
#include "direct_schur.cuh"
#include "structure.hpp"
#include "structure_device.cuh"

namespace gb {

// scratch device memory of an export call: freed on every return path
struct Scratch {
  void *p = nullptr;
  ~Scratch() { if (p) cudaFree(p); }
  template <typename X> X *as() const { return (X *)p; }
};

struct ProblemBase {
  gb_context *ctx = nullptr;
  HostStructure hs;
  virtual ~ProblemBase() {}
  virtual int init() = 0;
  virtual int set_observations(const void *) = 0;
  virtual int stage_observations_async(const void *, int) = 0;
  virtual int commit_observations(int) = 0;
  virtual int set_vertices(const void *, const void *) = 0;
  virtual int set_observations_device(const void *) = 0;
  virtual int set_vertices_device(const void *, const void *) = 0;
  virtual int get_vertices_device(void *, void *) = 0;
  virtual int import_linearization(const void *, const void *, const void *) = 0;
  virtual int solve_device(const gb_pcg_options *, void *, gb_solve_info *) = 0;
  virtual int set_factor(gb_factor_fn, void *) = 0;
  virtual int set_loss(int, double) = 0;
  virtual int set_fixed(const uint8_t *, const uint8_t *) = 0;
  virtual int set_precision(const void *) = 0;
  virtual int get_vertices(void *, void *) = 0;
  virtual int linearize(double *) = 0;
  virtual int compute_cost(double *) = 0;
  virtual int get_gradient(void *) = 0;
  virtual int get_scales(void *) = 0;
  virtual int get_residuals(void *) = 0;
  virtual int get_jacobians(double *, double *) = 0;
  virtual int hessian_values(void *) = 0;
  virtual int set_damping(double, int) = 0;
  virtual int solve(const gb_pcg_options *, void *, gb_solve_info *) = 0;
  virtual int get_schur_rhs(void *) = 0;
  virtual int get_schur_diagonal(void *) = 0;
  virtual int schur_multiply(const void *, void *) = 0;
  virtual int schur_structure(int64_t *, int64_t *, int64_t *) = 0;
  virtual int schur_values(void *) = 0;
  virtual int schur_csc(int32_t *, int32_t *, void *, int64_t *) = 0;
  virtual int try_step(double *, double *) = 0;
  virtual int revert_step() = 0;
  virtual int lm(const gb_lm_options *, gb_lm_result *, double *) = 0;
  virtual int time_stage(int, int, double *) = 0;
  virtual int structure_array(int, void *, int64_t *) = 0;
  virtual int64_t device_bytes() const = 0;
  virtual int exchange_mode() const = 0; // 0 single rank, 1 NCCL all-reduce, 2 peer-memory exchange (p2p.cuh)
};

template <typename T, typename S> struct Problem : ProblemBase {
  using T2 = typename V2<T>::type;
  using S2 = typename V2<S>::type;
  DevStruct ts{};
  std::vector<void *> allocs;
  int64_t bytes = 0;
  // state
  T *cams = nullptr, *pts = nullptr, *cams_bak = nullptr, *pts_bak = nullptr;
  T *camx = nullptr; // [Nc][CAMX] per-camera precomputed model terms (k_cam_precompute)
  T2 *obs = nullptr, *res = nullptr, *obs_stage = nullptr; // obs/res per storage slot; stage in caller order
  T2 *obs_stage2[2] = {nullptr, nullptr};                   // double-buffered staging of the asynchronous upload
  cudaEvent_t ev_stage[2] = {nullptr, nullptr};
  cudaEvent_t ev_slot_free = nullptr;
  bool staged[2] = {false, false};
  const int64_t *d_perm = nullptr;                          // sorted position -> caller index (null = identity)
  S2 *J = nullptr; // tile-major [ntiles][12][256]
  T *Cg = nullptr, *part18 = nullptr, *part54 = nullptr, *part9 = nullptr, *sums54 = nullptr;
  T *dot_part = nullptr, *rz_part = nullptr;
  // the PCG solve is ONE persistent cooperative kernel (k_pcg_solve, pcg_solve.cuh): one CTA per SM
  int solve_grid = 0;
  T *pbuf = nullptr, *st_dot = nullptr, *cta_red = nullptr;
  unsigned int *work_counter = nullptr;
  unsigned long long *d_timing = nullptr, *h_timing = nullptr; // per-iteration phase stamps of CTA 0 (profile_product)
  int timing_cap = 0;
  // explicit Schur complement as a solve mode (explicit_schur.cuh): structure and buffers are built on first use
  HostStructure::ExplicitSchur xs_host;
  ExplicitDev xs_dev{};
  bool xs_ready = false;
  T *xs_vals = nullptr, *xs_p = nullptr, *xs_red = nullptr;
  int xs_grid = 0;
  int last_schur_mode = GB_SCHUR_IMPLICIT;
  // direct solve of the reduced system (direct_schur.cuh)
  T *dn_A = nullptr;
  int *dn_fail = nullptr;
  int dn_grid = 0;
  // NCCL fallback (no peer memory between the ranks): product + reduction + all-reduce + cooperative update per iteration
  int64_t pcg_guess = 1 << 20; // iterations the previous solve executed
  T *diagB = nullptr, *gc = nullptr, *scale = nullptr /*[9Nc+3Np]*/, *b = nullptr /*[9Nc+3Np]*/;
  T *W = nullptr, *h = nullptr;
  T *Sdiag = nullptr, *Minv = nullptr, *bS = nullptr, *dterm = nullptr;
  T *x = nullptr, *r = nullptr, *z = nullptr, *pv = nullptr, *Ap = nullptr, *Ap_raw = nullptr, *xbak = nullptr, *xs = nullptr;
  T *delta = nullptr; // [9Nc+3Np] scaled-space step
  double *cost_part = nullptr, *rho_part = nullptr, *scalars = nullptr; // scalars: [0]=cost [1]=rho points [2]=rho cams
  PcgState<T> *pcg_state = nullptr;
  int *done_flag = nullptr;
  int pcg_state_cap = 0;
  // host mirrors
  double *h_scalars = nullptr;
  PcgState<T> *h_state = nullptr;
  // flags
  bool camx_valid = false; // camx holds the precomputed camera-model terms of the current cameras
  bool have_obs = false, have_vertices = false, linearized = false, prepared = false, solved = false, stepped = false;
  bool scale_on = true;
  // Low-precision Jacobian storage (S = bf16): the stored Jacobians are the reference's scaled, twice-rounded J~
  // (k_linearize<RESCALE>), the algebra downstream runs with D = I, and only the parameter update uses the true Jacobi
  // scales, kept in scale_true.
  static constexpr bool prescaled = IsLowPrecision<S>::value;
  T *scale_true = nullptr;
  // point sums of the PCG iterations kept by k_pcg_solve for the Jacobian-free back-substitution (k_backsubst_points)
  static constexpr int TK_MAX = 16; // solves with more iterations stream the Jacobians in the back-substitution instead
  T *tk_buf = nullptr, *tk_alpha = nullptr;
  int tk_alloc = 0;   // iterations tk_buf has room for
  int tk_iters = 0;   // > 0: the last solve left its point sums for this many iterations (0: use k_backsubst_tiles)
  int rho_pt_n = 0;   // entries of rho_part the last back-substitution wrote
  // fixed vertices (gb_set_fixed): device masks, null when nothing is fixed
  unsigned char *fixed_c = nullptr, *fixed_p = nullptr;
  bool any_fixed = false;
  int alg_scale() const { return (scale_on && !prescaled) ? 1 : 0; }
  const T *apply_scale() const { return prescaled ? scale_true : scale; }
  T mu = T(1e-4);
  int use_identity = 0;
  int ncamblocks = 0;
  gb_pcg_options last_pcg{10, 1.0, 5.0, 0, 0};
  int64_t dimc = 0, dimH = 0;
  cudaEvent_t ev[10]; // [8], [9]: around the re-linearisation of an accepted step (read at the next synchronisation)
  double last_chi2 = 0.0;
  bool profiling = false;
  // user-defined factor: evaluated by the caller's kernel into caller-order buffers (gb_set_factor)
  gb_factor_fn ext_fn = nullptr;
  void *ext_user = nullptr;
  T *ext_r = nullptr, *ext_Jc = nullptr, *ext_Jp = nullptr;
  int32_t *d_slot_src = nullptr, *d_ci_caller = nullptr, *d_pi_caller = nullptr;
  const T2 *obs_caller = nullptr; // the observations in the caller's order as last uploaded
  ExtFactor ex{nullptr, nullptr, nullptr, nullptr, nullptr};
  // loss and per-factor precision matrices (whitening in k_linearize / k_cost_tiles)
  T *Pu = nullptr;
  Robust rb{nullptr, 0, 0.0};
  // multi-GPU exchange over peer memory (p2p.cuh); NCCL is the bootstrap and the fallback
  P2P pp{};
  bool p2p_on = false;
  int *h_p2p_err = nullptr;
  void *p2p_area = nullptr;                  // own receive area + flag words (one cudaMalloc, exported by cudaIpc)
  void *p2p_peer[P2P_MAX_RANKS] = {nullptr}; // peers' areas as mapped here
  // full-system PCG (solver/pcg.hpp): work vectors of the whole Hessian dimension, allocated on first use
  T *f_x = nullptr, *f_r = nullptr, *f_z = nullptr, *f_p = nullptr, *f_v2 = nullptr, *f_xbak = nullptr;
  T *f_upw = nullptr, *f_outp = nullptr, *f_zero = nullptr, *f_Bfull = nullptr, *f_MinvF = nullptr, *f_part = nullptr;
  FullState<T> *f_state = nullptr, *h_fstate = nullptr; // PCG scalars of the full-system solver (device) and their pinned host copy
  bool full_info_pending = false;
  double *f_rho = nullptr;
  bool full_alloc = false, full_lin_valid = false, solved_full = false;
  int full_blocks = 0;
  gb_solve_info full_info{};

  ~Problem() override {
    if (ctx) cudaSetDevice(ctx->device);
    if (p2p_on) {
      // nobody may unmap or free while a peer can still be pushing: rendezvous first
      allreduce(scalars, 1, true);
      cudaStreamSynchronize(ctx->stream);
      for (int r = 0; r < ctx->nranks; r++)
        if (r != ctx->rank && p2p_peer[r]) cudaIpcCloseMemHandle(p2p_peer[r]);
      allreduce(scalars, 1, true);
      cudaStreamSynchronize(ctx->stream);
    }
    if (p2p_area) cudaFree(p2p_area);
    for (void *p : allocs) cudaFree(p);
    if (h_scalars) cudaFreeHost(h_scalars);
    if (h_state) cudaFreeHost(h_state);
    if (h_p2p_err) cudaFreeHost(h_p2p_err);
    if (h_fstate) cudaFreeHost(h_fstate);
    for (auto &e : ev) if (e) cudaEventDestroy(e);
    if (h_timing) cudaFreeHost(h_timing);
    for (auto &e : ev_stage) if (e) cudaEventDestroy(e);
    if (ev_slot_free) cudaEventDestroy(ev_slot_free);
  }
  int64_t device_bytes() const override { return bytes; }
  int exchange_mode() const override { return ctx->nranks <= 1 ? 0 : (p2p_on ? 2 : 1); }

  template <typename X> int dalloc(X *&p, size_t n) {
    void *q = nullptr;
    const size_t sz = std::max<size_t>(n, 1) * sizeof(X);
    GB_CUDA(ctx, cudaMalloc(&q, sz));
    GB_CUDA(ctx, cudaMemsetAsync(q, 0, sz, ctx->stream));
    allocs.push_back(q);
    bytes += (int64_t)sz;
    p = (X *)q;
    return GB_OK;
  }
  template <typename X> int upload(const X *&dst, const std::vector<X> &v) {
    X *p = nullptr;
    GB_TRY(dalloc(p, v.size()));
    GB_CUDA(ctx, cudaMemcpyAsync(p, v.data(), v.size() * sizeof(X), cudaMemcpyHostToDevice, ctx->stream));
    dst = p;
    return GB_OK;
  }

  // device copies of the structure tables (gb_problem_structure_array): the GPU-built arrays against the host view
  int structure_array(int which, void *out, int64_t *count) override {
    const void *src = nullptr;
    size_t bytes = 0;
    switch (which) {
      case 10: src = ts.slot_of_obs; bytes = (size_t)ts.M * 4; *count = ts.M; break;
      case 13: src = ts.ometa; bytes = (size_t)ts.Mstore * 4; *count = ts.Mstore; break;
      case 17: src = ts.trec; bytes = (size_t)ts.ntiles * REC_BYTES; *count = (int64_t)bytes; break;
      case 18: src = ts.tile_cam; bytes = (size_t)ts.Mstore * 4; *count = ts.Mstore; break;
      case 19: src = ts.cm_slot; bytes = (size_t)ts.M * 4; *count = ts.M; break;
      case 20: src = ts.cm_pt; bytes = (size_t)ts.M * 4; *count = ts.M; break;
      default: return ctx->fail(GB_ERR_INVALID, "gb_problem_structure_array: unknown array %d", which);
    }
    if (out) {
      GB_CUDA(ctx, cudaMemcpyAsync(out, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
      GB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return GB_OK;
  }

  int init() override {
    GB_CUDA(ctx, cudaSetDevice(ctx->device));
    for (auto &e : ev) e = nullptr;
    for (auto &e : ev) GB_CUDA(ctx, cudaEventCreate(&e));
    const int64_t M = hs.M, Nc = hs.Nc, Np = hs.Np;
    dimc = 9 * Nc;
    dimH = 9 * Nc + 3 * Np;
    ts.M = M;
    ts.Mstore = hs.Mstore;
    ts.Nc = hs.Nc; ts.Np = hs.Np; ts.ntiles = hs.ntiles(); ts.nst = hs.nst(); ts.nrows = hs.nrows(); ts.pad = 0;
    GB_TRY(upload(ts.tmeta, hs.tmeta));
    GB_TRY(upload(ts.st_tile, hs.st_tile));
    GB_TRY(upload(ts.st_row, hs.st_row));
    GB_TRY(upload(ts.row_cam, hs.row_cam));
    GB_TRY(upload(ts.cam_row_ptr, hs.cam_row_ptr));
    GB_TRY(upload(ts.row_out, hs.row_out));
    GB_TRY(upload(ts.cta_st, hs.cta_st));
    GB_TRY(upload(ts.strec, hs.strec));
    ts.ncta = hs.ncta(); ts.pad2 = 0;
    GB_TRY(upload(ts.ch_ptr, hs.ch_ptr));
    GB_TRY(upload(ts.cam_ch_ptr, hs.cam_ch_ptr));
    ts.nchunks = hs.nchunks(); ts.pad3 = 0;
    GB_TRY(upload(ts.cam_idx, hs.cam_idx));
    GB_TRY(upload(ts.pt_idx, hs.pt_idx));
    GB_TRY(upload(ts.pptr, hs.pptr));
    // long tracks: fragment tables and the two-level sum buffers (structure.hpp, kernels.cuh "long tracks")
    ts.tk = nullptr; ts.tk_alpha = nullptr; ts.tk_cap = 0; ts.pad4 = 0;
    ts.nfrag = hs.nfrag(); ts.nheavy = hs.nheavy();
    GB_TRY(upload(ts.hv_pt, hs.hv_pt)); GB_TRY(upload(ts.hv_ptr, hs.hv_ptr));
    {
      T *fp = nullptr, *ft = nullptr;
      GB_TRY(dalloc(fp, 9 * (size_t)ts.nfrag)); GB_TRY(dalloc(ft, 3 * (size_t)ts.nfrag));
      ts.frag_part = fp; ts.frag_t = ft;
    }
    if (hs.tables_on_device) {
      // the observation-sized tables are built here, on the device (structure_device.cuh); the host made the cuts only
      const int32_t *d_tile_obs = nullptr, *d_tile_pt = nullptr, *d_tile_st = nullptr;
      GB_TRY(upload(d_tile_obs, hs.tile_obs));
      GB_TRY(upload(d_tile_pt, hs.tile_pt));
      GB_TRY(upload(d_tile_st, hs.tile_st));
      uint32_t *d_ometa = nullptr;
      unsigned char *d_trec = nullptr;
      int32_t *d_tile_cam = nullptr, *d_slot = nullptr, *d_cm_slot = nullptr, *d_cm_pt = nullptr;
      GB_TRY(dalloc(d_ometa, hs.Mstore)); GB_TRY(dalloc(d_trec, (size_t)ts.ntiles * REC_BYTES));
      GB_TRY(dalloc(d_tile_cam, hs.Mstore)); GB_TRY(dalloc(d_slot, M));
      GB_TRY(dalloc(d_cm_slot, M)); GB_TRY(dalloc(d_cm_pt, M));
      k_build_tile_tables<<<ts.ntiles, TILE, 0, ctx->stream>>>(ts.ntiles, ts.cam_idx, ts.pt_idx, ts.pptr, d_tile_obs, d_tile_pt, d_tile_st,
                                                                ts.st_row, ts.row_cam, ts.tmeta, d_ometa, d_trec, d_tile_cam, d_slot);
      GB_LAUNCH(ctx);
      {
        Scratch k1, k2, k3;
        GB_CUDA(ctx, cudaMalloc(&k1.p, (size_t)M * 4)); GB_CUDA(ctx, cudaMalloc(&k2.p, (size_t)M * 4)); GB_CUDA(ctx, cudaMalloc(&k3.p, (size_t)M * 4));
        GB_CUDA(ctx, camera_major_order(M, hs.Nc, ts.cam_idx, k1.as<int32_t>(), k2.as<int32_t>(), k3.as<int32_t>(), ctx->stream));
        k_camera_major<<<(unsigned)((M + 255) / 256), 256, 0, ctx->stream>>>(M, k3.as<int32_t>(), d_slot, ts.pt_idx, d_cm_slot, d_cm_pt);
        GB_LAUNCH(ctx);
        GB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      }
      ts.ometa = d_ometa; ts.trec = d_trec; ts.tile_cam = d_tile_cam; ts.slot_of_obs = d_slot; ts.cm_slot = d_cm_slot; ts.cm_pt = d_cm_pt;
      // host-side exports permute through slot_of_obs: keep a copy (20 MB at Venice)
      hs.slot_of_obs.resize((size_t)M);
      GB_CUDA(ctx, cudaMemcpyAsync(hs.slot_of_obs.data(), d_slot, (size_t)M * 4, cudaMemcpyDeviceToHost, ctx->stream));
      GB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    } else {
      GB_TRY(upload(ts.ometa, hs.ometa));
      GB_TRY(upload(ts.trec, hs.trec));
      GB_TRY(upload(ts.tile_cam, hs.tile_cam));
      GB_TRY(upload(ts.cm_slot, hs.cm_slot));
      GB_TRY(upload(ts.cm_pt, hs.cm_pt));
      GB_TRY(upload(ts.slot_of_obs, hs.slot_of_obs));
    }
    GB_TRY(dalloc(cams, Nc * CAM_STRIDE)); GB_TRY(dalloc(cams_bak, Nc * CAM_STRIDE)); GB_TRY(dalloc(camx, Nc * CAMX));
    GB_TRY(dalloc(pts, 3 * Np)); GB_TRY(dalloc(pts_bak, 3 * Np));
    GB_TRY(dalloc(obs, hs.Mstore)); GB_TRY(dalloc(res, hs.Mstore)); GB_TRY(dalloc(obs_stage, M));
    if (!hs.identity_perm) GB_TRY(upload(d_perm, hs.perm));
    GB_TRY(dalloc(J, (size_t)NPLANES * hs.Mstore)); // zero-filled: padding slots stay zero for ever
    GB_TRY(dalloc(Cg, 9 * Np));
    GB_TRY(dalloc(part18, (size_t)ts.nrows * 18)); GB_TRY(dalloc(part54, (size_t)std::max(ts.nchunks, 1) * 54));
    GB_TRY(dalloc(part9, (size_t)ts.nrows * 9)); GB_TRY(dalloc(sums54, Nc * 54));
    GB_TRY(dalloc(dot_part, Nc)); GB_TRY(dalloc(rz_part, Nc));
    {
      int per_sm = 0, per_sm_solve = 0, sms = 0, coop = 0;
      GB_CUDA(ctx, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
      GB_CUDA(ctx, cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device));
      // the super-tile kernels keep their camera accumulator rows in (opt-in sized) dynamic shared memory
      GB_CUDA(ctx, cudaFuncSetAttribute(k_linearize<T, S, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_lin_bytes<T>()));
      GB_CUDA(ctx, cudaFuncSetAttribute(k_linearize<T, S, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_lin_bytes<T>()));
      GB_CUDA(ctx, cudaFuncSetAttribute(k_linearize<T, S, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_lin_bytes<T>()));
      GB_CUDA(ctx, cudaFuncSetAttribute(k_schur_product2<T, S, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        SchurSmem2<T, S>::TOTAL));
      GB_CUDA(ctx, cudaFuncSetAttribute(k_schur_product2<T, S, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        SchurSmem2<T, S>::TOTAL));
      GB_CUDA(ctx, cudaFuncSetAttribute(k_pcg_solve<T, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, SolveSmem<T, S>::TOTAL));
      GB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pcg_update<T>, 288, 0));
      GB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_solve, k_pcg_solve<T, S>, SolveSmem<T, S>::THREADS,
                                                                 SolveSmem<T, S>::TOTAL));
      if (!coop || per_sm_solve < 1 || (int64_t)per_sm * sms < (Nc + PCG_CAMS - 1) / PCG_CAMS)
        return ctx->fail(GB_ERR_UNSUPPORTED, "the device cannot co-schedule the PCG kernels (cooperative launch)");
      solve_grid = sms; // the same on every rank: the per-CTA exchange pairs CTA b with CTA b of the peers
      GB_TRY(dalloc(cta_red, 2 * (size_t)solve_grid));
      GB_TRY(dalloc(pbuf, 2 * (size_t)dimc));
      GB_TRY(dalloc(st_dot, (size_t)ts.nst));
      GB_TRY(dalloc(work_counter, 1));
    }
    GB_TRY(dalloc(diagB, 2 * dimc)); // diag(B) | g_c contiguous: one exchange for both on the multi-GPU path
    gc = diagB + dimc;
    GB_TRY(dalloc(scale, dimH)); GB_TRY(dalloc(b, dimH)); GB_TRY(dalloc(delta, dimH));
    if (prescaled) GB_TRY(dalloc(scale_true, dimH));
    GB_TRY(dalloc(W, (size_t)WST<T>::value * Np + 8)); GB_TRY(dalloc(h, (size_t)HST * Np + 8));
    GB_TRY(dalloc(Sdiag, Nc * 81)); GB_TRY(dalloc(Minv, Nc * 81));
    GB_TRY(dalloc(bS, dimc)); GB_TRY(dalloc(dterm, dimc));
    GB_TRY(dalloc(x, dimc)); GB_TRY(dalloc(r, dimc)); GB_TRY(dalloc(z, dimc)); GB_TRY(dalloc(pv, dimc));
    GB_TRY(dalloc(Ap, dimc)); GB_TRY(dalloc(Ap_raw, dimc)); GB_TRY(dalloc(xbak, dimc));
    GB_TRY(dalloc(xs, Nc * CAM_STRIDE));
    ncamblocks = (int)((dimc + 255) / 256);
    GB_TRY(dalloc(cost_part, ts.ntiles)); GB_TRY(dalloc(rho_part, ts.ntiles + ncamblocks));
    GB_TRY(dalloc(scalars, 8));
    GB_TRY(dalloc(done_flag, 1));
    pcg_state_cap = 2 * 64 + 3;
    GB_TRY(dalloc(pcg_state, pcg_state_cap));
    GB_CUDA(ctx, cudaMallocHost((void **)&h_scalars, 8 * sizeof(double)));
    GB_CUDA(ctx, cudaMallocHost((void **)&h_state, sizeof(PcgState<T>)));
    GB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->nranks > 1) GB_TRY(p2p_setup());
    return GB_OK;
  }

  // Peer-memory exchange: one receive area per rank ([2 halves][nranks][slot] + flag words), exported with cudaIpc;
  // the handles travel through an NCCL all-gather.  Any failure (no peer access, IPC unavailable) leaves p2p_on false
  // on ALL ranks (agreed through an all-reduce) and the collectives stay on NCCL.  GB_P2P=0 disables it.
  int p2p_setup() {
    const int n = ctx->nranks, me = ctx->rank;
    const char *env = getenv("GB_P2P");
    int fail = (env && env[0] == '0') || n > P2P_MAX_RANKS || !g_nccl.AllGather ? 1 : 0;
    const size_t slot = (((size_t)54 * ts.Nc * sizeof(T) + 1024) + 255) / 256 * 256;
    // [2 halves][n slots] | flag words (256 B) | magic (256 B) | LL area of k_pcg_solve [2 halves][n slots]: every 32-bit
    // half of a value in an 8-byte word with its epoch (9 Nc values + the dot scalar per slot)
    const size_t half = slot * n, flag_off = 2 * half, magic_off = flag_off + 256, ll_off = magic_off + 256;
    const size_t ll_slot = (((size_t)9 * ts.Nc + 8) * (sizeof(T) / 4) * 8 + 255) / 256 * 256, ll_half = ll_slot * n;
    const size_t area_end = ll_off + 2 * ll_half;
    const size_t total = std::max<size_t>((area_end + (1 << 21) - 1) >> 21 << 21, (size_t)4 << 20); // own VA range
    struct Pack { cudaIpcMemHandle_t h; unsigned long long magic; };
    std::vector<Pack> packs(n);
    unsigned char *hbuf = nullptr;
    unsigned int *d_counter = nullptr;
    unsigned long long *d_seq = nullptr;
    GB_TRY(dalloc(hbuf, sizeof(Pack) * n));
    GB_TRY(dalloc(d_counter, 1)); GB_TRY(dalloc(d_seq, 1));
    // the time-out flag lives in pinned host memory (device-visible through UVA): the host reads it without a copy
    GB_CUDA(ctx, cudaMallocHost((void **)&h_p2p_err, sizeof(int)));
    *h_p2p_err = 0;
    const unsigned long long magic = 0x67623230ull * 1000003ull + (unsigned long long)me * 7919ull + (unsigned long long)(uintptr_t)this;
    if (!fail) {
      if (cudaMalloc(&p2p_area, total) != cudaSuccess) { cudaGetLastError(); p2p_area = nullptr; fail = 1; }
    }
    if (!fail) {
      GB_CUDA(ctx, cudaMemsetAsync(p2p_area, 0, total, ctx->stream));
      GB_CUDA(ctx, cudaMemcpyAsync((unsigned char *)p2p_area + magic_off, &magic, 8, cudaMemcpyHostToDevice, ctx->stream));
      GB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      Pack mine{};
      if (cudaIpcGetMemHandle(&mine.h, p2p_area) != cudaSuccess) { cudaGetLastError(); fail = 1; }
      mine.magic = magic;
      GB_CUDA(ctx, cudaMemcpyAsync(hbuf + sizeof(Pack) * me, &mine, sizeof(Pack), cudaMemcpyHostToDevice, ctx->stream));
    }
    if (g_nccl.AllGather) {
      const int rc = g_nccl.AllGather(hbuf + sizeof(Pack) * me, hbuf, sizeof(Pack), NCCL_INT8, ctx->comm, ctx->stream);
      if (rc != 0) return ctx->fail(GB_ERR_NCCL, "ncclAllGather failed");
      GB_CUDA(ctx, cudaMemcpyAsync(packs.data(), hbuf, sizeof(Pack) * n, cudaMemcpyDeviceToHost, ctx->stream));
      GB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    if (!fail) {
      for (int r = 0; r < n && !fail; r++) {
        if (r == me) { p2p_peer[r] = p2p_area; continue; }
        void *q = nullptr;
        if (cudaIpcOpenMemHandle(&q, packs[r].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); fail = 1; break; }
        p2p_peer[r] = q;
        unsigned long long seen = 0; // the mapping must start at the peer's allocation base
        if (cudaMemcpy(&seen, (unsigned char *)q + magic_off, 8, cudaMemcpyDeviceToHost) != cudaSuccess) { cudaGetLastError(); fail = 1; }
        else if (seen != packs[r].magic) fail = 1;
      }
    }
    // agree: everybody or nobody
    int *d_fail = nullptr;
    GB_TRY(dalloc(d_fail, 1));
    GB_CUDA(ctx, cudaMemcpyAsync(d_fail, &fail, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    const int rc = g_nccl.AllReduce(d_fail, d_fail, 1, NCCL_INT32, NCCL_SUM, ctx->comm, ctx->stream);
    if (rc != 0) return ctx->fail(GB_ERR_NCCL, "ncclAllReduce failed");
    int total_fail = 0;
    GB_CUDA(ctx, cudaMemcpyAsync(&total_fail, d_fail, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    GB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (total_fail != 0) {
      for (int r = 0; r < n; r++)
        if (r != me && p2p_peer[r]) { cudaIpcCloseMemHandle(p2p_peer[r]); p2p_peer[r] = nullptr; }
      return GB_OK; // NCCL path
    }
    pp.nranks = n; pp.rank = me;
    for (int r = 0; r < n; r++) {
      pp.recv[r] = (unsigned char *)p2p_peer[r];
      pp.flags[r] = (unsigned long long *)((unsigned char *)p2p_peer[r] + flag_off);
    }
    pp.half_bytes = half; pp.slot_bytes = slot;
    pp.ll_off = ll_off; pp.ll_half_bytes = ll_half; pp.ll_slot_bytes = ll_slot;
    pp.counter = d_counter; pp.seq = d_seq; pp.error = h_p2p_err;
    {
      // ranks are only loosely in step on the host (structure builds, uploads): a consumer waits this long for a peer
      // before the call fails with GB_ERR_NCCL (every multi-rank entry point checks the flag after its synchronise)
      const char *envt = getenv("GB_P2P_TIMEOUT_S");
      double sec = envt ? atof(envt) : 120.0;
      if (!(sec > 0.0)) sec = 120.0;
      pp.timeout_ns = (unsigned long long)(sec * 1e9);
    }
    p2p_on = true;
    return GB_OK;
  }
  int p2p_check() { // after a synchronise: did any wait time out?
    if (!p2p_on) return GB_OK;
    if (*(volatile int *)h_p2p_err) return ctx->fail(GB_ERR_NCCL, "peer-memory exchange timed out: a rank did not arrive");
    return GB_OK;
  }

  // device -> pinned host without the copy engine (k_store_host)
  int store_host(void *host_dst, const void *src, size_t bytes) {
    k_store_host<<<1, 32, 0, ctx->stream>>>((uint32_t *)host_dst, (const uint32_t *)src, (int)(bytes / 4));
    GB_LAUNCH(ctx);
    return GB_OK;
  }
  int ensure_state_cap(int64_t max_iter) {
    const int need = (int)(2 * max_iter + 3);
    if (need > pcg_state_cap) {
      GB_TRY(dalloc(pcg_state, need));
      pcg_state_cap = need;
    }
    return GB_OK;
  }

  // ---- collectives -----------------------------------------------------------------------------------
  int allreduce(void *buf, size_t count, bool is_double) {
    if (ctx->nranks <= 1) return GB_OK;
    const int rc = g_nccl.AllReduce(buf, buf, count, is_double ? NCCL_FLOAT64 : NCCL_FLOAT32, NCCL_SUM, ctx->comm, ctx->stream);
    if (rc != 0) return ctx->fail(GB_ERR_NCCL, "ncclAllReduce failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
    return GB_OK;
  }
  // sum over ranks, in place.  Peer-memory path: exchange_push stores the buffer into every rank's receive slot,
  // exchange_sum waits for the peers and adds the slots in rank order; independent local kernels may be enqueued between
  // the two so that they run while the peers' data is in flight.  NCCL fallback: the all-reduce happens in exchange_sum.
  template <typename X> bool exchange_p2p(size_t count) const { return p2p_on && count * sizeof(X) <= pp.slot_bytes; }
  template <typename X> int exchange_push(X *buf, size_t count) {
    if (ctx->nranks <= 1 || !exchange_p2p<X>(count)) return GB_OK;
    const int grid = (int)std::min<size_t>((count + 255) / 256, 4 * 148);
    k_p2p_push<X><<<grid, 256, 0, ctx->stream>>>(pp, buf, (long long)count);
    GB_LAUNCH(ctx);
    return GB_OK;
  }
  template <typename X> int exchange_sum(X *buf, size_t count) {
    if (ctx->nranks <= 1) return GB_OK;
    if (!exchange_p2p<X>(count)) return allreduce(buf, count, sizeof(X) == 8);
    const int grid = (int)std::min<size_t>((count + 255) / 256, 4 * 148);
    k_p2p_sum<X><<<grid, 256, 0, ctx->stream>>>(pp, buf, (long long)count);
    GB_LAUNCH(ctx);
    return GB_OK;
  }
  template <typename X> int exchange(X *buf, size_t count) {
    GB_TRY(exchange_push<X>(buf, count));
    return exchange_sum<X>(buf, count);
  }
  int allreduce_T(T *buf, size_t count) { return exchange<T>(buf, count); }

  // ---- IO ----------------------------------------------------------------------------------------------
  int set_observations(const void *o) override {
    // one H2D copy in the caller's order, then a device scatter into the tile-padded storage slots
    GB_CUDA(ctx, cudaMemcpyAsync(obs_stage, o, 2 * hs.M * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    k_scatter_slots<T2><<<(unsigned)((hs.M + 255) / 256), 256, 0, ctx->stream>>>(hs.M, ts.slot_of_obs, d_perm, obs_stage, obs);
    GB_LAUNCH(ctx);
    obs_caller = obs_stage;
    GB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    have_obs = true;
    linearized = prepared = solved = solved_full = full_lin_valid = stepped = false;
    return GB_OK;
  }
  // Double-buffered observation upload on the context's copy stream: the H2D copy of the NEXT batch overlaps whatever
  // the compute stream is doing (the LM iteration on the current batch).  stage -> returns at once; commit -> the compute
  // stream waits for that slot's copy and scatters it into the tile-padded storage.  The host buffer must be pinned
  // and stay valid until the commit.
  int stage_observations_async(const void *o, int slot) override {
    if (slot < 0 || slot > 1) return ctx->fail(GB_ERR_INVALID, "staging slot must be 0 or 1");
    if (!obs_stage2[slot]) {
      GB_TRY(dalloc(obs_stage2[slot], hs.M));
      GB_CUDA(ctx, cudaEventCreateWithFlags(&ev_stage[slot], cudaEventDisableTiming));
      if (!ev_slot_free) GB_CUDA(ctx, cudaEventCreateWithFlags(&ev_slot_free, cudaEventDisableTiming));
      GB_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); // dalloc's memset is on the compute stream
    }
    // the slot may still be read by work already enqueued on the compute stream (the scatter of its last commit, a
    // user factor reading obs_caller): the copy starts only after that work, not after work enqueued later
    GB_CUDA(ctx, cudaEventRecord(ev_slot_free, ctx->stream));
    GB_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ev_slot_free, 0));
    GB_CUDA(ctx, cudaMemcpyAsync(obs_stage2[slot], o, 2 * hs.M * sizeof(T), cudaMemcpyHostToDevice, ctx->copy_stream));
    GB_CUDA(ctx, cudaEventRecord(ev_stage[slot], ctx->copy_stream));
    staged[slot] = true;
    return GB_OK;
  }
  int commit_observations(int slot) override {
    if (slot < 0 || slot > 1 || !staged[slot]) return ctx->fail(GB_ERR_INVALID, "gb_commit_observations: nothing staged in that slot");
    GB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ev_stage[slot], 0));
    k_scatter_slots<T2><<<(unsigned)((hs.M + 255) / 256), 256, 0, ctx->stream>>>(hs.M, ts.slot_of_obs, d_perm, obs_stage2[slot], obs);
    GB_LAUNCH(ctx);
    obs_caller = obs_stage2[slot]; // stays untouched until this slot is staged again
    staged[slot] = false;
    have_obs = true;
    linearized = prepared = solved = solved_full = stepped = false;
    return launch_check();
  }
  int set_vertices(const void *c, const void *p) override {
    // 9 -> 10 padded rows: a strided 2-D copy, no host staging
    GB_CUDA(ctx, cudaMemcpy2DAsync(cams, CAM_STRIDE * sizeof(T), c, 9 * sizeof(T), 9 * sizeof(T), hs.Nc,
                                   cudaMemcpyHostToDevice, ctx->stream));
    GB_CUDA(ctx, cudaMemcpyAsync(pts, p, 3 * (size_t)hs.Np * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    // Leave the stream on the compute engine: the first kernel after a copy waits on a semaphore the driver appends to
    // the copy engine's queue at that moment - behind any large upload another stream has queued there meanwhile
    // (measured: every step waited the full 1.5 ms of the overlapped observation upload, scripts/e2e_probe.py).
    k_cam_precompute<T><<<(ts.Nc + 127) / 128, 128, 0, ctx->stream>>>(ts.Nc, cams, camx);
    GB_LAUNCH(ctx);
    camx_valid = true;
    GB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    have_vertices = true;
    linearized = prepared = solved = solved_full = full_lin_valid = stepped = false;
    return GB_OK;
  }
  // storage slot -> caller's factor index, camera / point index per caller factor: what a caller-order buffer needs
  int ensure_caller_maps() {
    if (d_slot_src) return GB_OK;
    std::vector<int32_t> src((size_t)hs.Mstore, -1), ci((size_t)hs.M), pi((size_t)hs.M);
    for (int64_t spos = 0; spos < hs.M; spos++) {
      const int64_t u = hs.identity_perm ? spos : hs.perm[spos];
      src[hs.slot_of_obs[spos]] = (int32_t)u;
      ci[u] = hs.cam_idx[spos];
      pi[u] = hs.pt_idx[spos];
    }
    const int32_t *a = nullptr, *b = nullptr, *c = nullptr;
    GB_TRY(upload(a, src)); GB_TRY(upload(b, ci)); GB_TRY(upload(c, pi));
    d_slot_src = const_cast<int32_t *>(a); d_ci_caller = const_cast<int32_t *>(b); d_pi_caller = const_cast<int32_t *>(c);
    GB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return GB_OK;
  }
  int set_observations_device(const void *o) override {
    k_scatter_slots<T2><<<(unsigned)((hs.M + 255) / 256), 256, 0, ctx->stream>>>(hs.M, ts.slot_of_obs, d_perm, (const T2 *)o, obs);
    GB_LAUNCH(ctx);
    obs_caller = (const T2 *)o; // the caller keeps the buffer alive while a user-defined factor may read it
    have_obs = true;
    linearized = prepared = solved = solved_full = full_lin_valid = stepped = false;
    return launch_check();
  }
  int set_vertices_device(const void *c, const void *p) override {
    GB_CUDA(ctx, cudaMemcpy2DAsync(cams, CAM_STRIDE * sizeof(T), c, 9 * sizeof(T), 9 * sizeof(T), hs.Nc,
                                   cudaMemcpyDeviceToDevice, ctx->stream));
    GB_CUDA(ctx, cudaMemcpyAsync(pts, p, 3 * (size_t)hs.Np * sizeof(T), cudaMemcpyDeviceToDevice, ctx->stream));
    camx_valid = false;
    have_vertices = true;
    linearized = prepared = solved = solved_full = full_lin_valid = stepped = false;
    return GB_OK;
  }
  int get_vertices_device(void *c, void *p) override {
    if (!have_vertices) return ctx->fail(GB_ERR_INVALID, "gb_get_vertices_device before the vertices were set");
    if (c)
      GB_CUDA(ctx, cudaMemcpy2DAsync(c, 9 * sizeof(T), cams, CAM_STRIDE * sizeof(T), 9 * sizeof(T), hs.Nc,
                                     cudaMemcpyDeviceToDevice, ctx->stream));
    if (p) GB_CUDA(ctx, cudaMemcpyAsync(p, pts, 3 * (size_t)hs.Np * sizeof(T), cudaMemcpyDeviceToDevice, ctx->stream));
    return GB_OK;
  }
  // Solver::update_values for a caller that linearises itself: the reference's residuals and SCALED Jacobians, in its
  // factor order, take the place of the built-in factor evaluation; everything downstream (assembly, Schur operator,
  // PCG) is the production path with scales = 1.
  int import_linearization(const void *rdev, const void *jc, const void *jp) override {
    if constexpr (!std::is_same<T, S>::value) {
      return ctx->fail(GB_ERR_UNSUPPORTED, "gb_import_linearization needs T == S");
    } else {
      GB_TRY(require(ctx->nranks == 1, "gb_import_linearization is single-rank"));
      GB_TRY(ensure_caller_maps());
      scale_on = false;
      have_obs = have_vertices = true; // the caller owns both; nothing here evaluates factors
      GB_TRY(enqueue_linearize(true, ExtFactor{rdev, jc, jp, d_slot_src, nullptr}));
      return GB_OK;
    }
  }
  int set_factor(gb_factor_fn fn, void *user) override {
    linearized = prepared = solved = solved_full = stepped = false;
    ext_fn = fn;
    ext_user = user;
    if (!fn || ext_r) return GB_OK;
    GB_TRY(dalloc(ext_r, 2 * (size_t)hs.M)); GB_TRY(dalloc(ext_Jc, 18 * (size_t)hs.M)); GB_TRY(dalloc(ext_Jp, 6 * (size_t)hs.M));
    GB_TRY(ensure_caller_maps());
    ex = ExtFactor{ext_r, ext_Jc, ext_Jp, d_slot_src, nullptr};
    return GB_OK;
  }
  // run the caller's factor kernel at the current vertices (with_jacobians = false: residuals only)
  int evaluate_external(bool with_jacobians) {
    if (!obs_caller) return ctx->fail(GB_ERR_INVALID, "user-defined factor: no observations uploaded");
    gb_factor_eval e{};
    e.num_observations = hs.M;
    e.cameras = cams; e.points = pts; e.observations = obs_caller;
    e.camera_index = d_ci_caller; e.point_index = d_pi_caller;
    e.residuals = ext_r;
    e.Jc = with_jacobians ? ext_Jc : nullptr;
    e.Jp = with_jacobians ? ext_Jp : nullptr;
    e.stream = (void *)ctx->stream;
    const int rc = ext_fn(&e, ext_user);
    if (rc != 0) return ctx->fail(GB_ERR_INVALID, "user-defined factor callback returned %d", rc);
    return launch_check();
  }
  // Replaces VertexDescriptor::set_fixed (vertex.hpp:254-266): fixed cameras / points keep their slot in every vector
  // (scale 0, gradient 0, step 0) instead of being left out of the Hessian ordering; every other unknown gets the same
  // numbers as in the reference's reduced system.
  int set_fixed(const uint8_t *fc, const uint8_t *fp) override {
    if (prescaled) return ctx->fail(GB_ERR_UNSUPPORTED, "fixed vertices with bf16 Jacobian storage: use the generic factor-graph path");
    auto install = [&](const uint8_t *src, int64_t n, unsigned char *&dst) -> int {
      bool any = false;
      if (src) for (int64_t i = 0; i < n; i++) any = any || src[i] != 0;
      if (!any) { dst = nullptr; return GB_OK; }
      unsigned char *buf = nullptr;
      GB_CUDA(ctx, cudaMalloc((void **)&buf, (size_t)n));
      allocs.push_back(buf);
      GB_CUDA(ctx, cudaMemcpyAsync(buf, src, (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
      GB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      dst = buf;
      return GB_OK;
    };
    GB_TRY(install(fc, ts.Nc, fixed_c));
    GB_TRY(install(fp, ts.Np, fixed_p));
    any_fixed = fixed_c != nullptr || fixed_p != nullptr;
    linearized = prepared = solved = solved_full = stepped = false;
    return GB_OK;
  }
  // Replaces the loss argument of add_factor (factor.hpp:373-412, loss.hpp): one loss for all factors.
  int set_loss(int kind, double delta) override {
    if (kind != 0 && kind != 1) return ctx->fail(GB_ERR_INVALID, "unknown loss %d", kind);
    if (kind == 1 && !(delta > 0.0)) return ctx->fail(GB_ERR_INVALID, "Huber delta must be positive");
    rb.loss_kind = kind;
    rb.delta = kind ? delta : 0.0;
    linearized = prepared = solved = solved_full = stepped = false;
    return GB_OK;
  }
  // Replaces the precision_matrix argument of add_factor: [n_obs][4] row-major 2x2 in the caller's factor order
  // (element type T), symmetric positive definite; NULL restores the identity.  Factored P = U^T U here.
  int set_precision(const void *Pin) override {
    linearized = prepared = solved = solved_full = stepped = false;
    if (!Pin) { rb.Pu = nullptr; return GB_OK; }
    const T *P = (const T *)Pin;
    std::vector<T> hu(3 * (size_t)hs.Mstore, T(0));
    for (int64_t spos = 0; spos < hs.M; spos++) {
      const int64_t u = hs.identity_perm ? spos : hs.perm[spos], sl = hs.slot_of_obs[spos];
      const double p00 = (double)P[4 * u], p01 = (double)P[4 * u + 1], p10 = (double)P[4 * u + 2], p11 = (double)P[4 * u + 3];
      const double tol = 1e-6 * std::max(std::fabs(p00), std::fabs(p11));
      if (!(p00 > 0.0) || std::fabs(p01 - p10) > tol || !(p00 * p11 - p01 * p10 > 0.0))
        return ctx->fail(GB_ERR_INVALID, "precision matrix of factor %ld is not symmetric positive definite", (long)u);
      const double u00 = std::sqrt(p00), u01 = p01 / u00, u11 = std::sqrt(p11 - u01 * u01);
      hu[3 * sl] = (T)u00; hu[3 * sl + 1] = (T)u01; hu[3 * sl + 2] = (T)u11;
    }
    if (!Pu) GB_TRY(dalloc(Pu, 3 * (size_t)hs.Mstore));
    GB_CUDA(ctx, cudaMemcpyAsync(Pu, hu.data(), hu.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    GB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    rb.Pu = Pu;
    return GB_OK;
  }
  int get_vertices(void *c, void *p) override {
    if (!have_vertices) return ctx->fail(GB_ERR_INVALID, "gb_get_vertices before gb_set_vertices");
    if (c)
      GB_CUDA(ctx, cudaMemcpy2DAsync(c, 9 * sizeof(T), cams, CAM_STRIDE * sizeof(T), 9 * sizeof(T), hs.Nc,
                                     cudaMemcpyDeviceToHost, ctx->stream));
    if (p) GB_CUDA(ctx, cudaMemcpyAsync(p, pts, 3 * (size_t)hs.Np * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
    GB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return GB_OK;
  }

  // ---- stages (asynchronous on the context stream) -----------------------------------------------------
  int launch_check() {
    GB_CUDA(ctx, cudaGetLastError());
    return GB_OK;
  }

  // long tracks: add the fragments' partial sums of `width` values per point into out[Np][width] (after k_linearize and
  // the full-system product) / form t_p = sum_o Jp^T Jc x_c of the long-track points from xs (before the Schur product
  // and the back-substitution)
  int enqueue_frag_sum(int width, T *out) {
    if (ts.nheavy == 0) return GB_OK;
    k_frag_sum<T><<<(ts.nheavy * width + 127) / 128, 128, 0, ctx->stream>>>(ts.nheavy, ts.hv_pt, ts.hv_ptr,
                                                                          reinterpret_cast<const T *>(ts.frag_part), width, out);
    GB_LAUNCH(ctx);
    return GB_OK;
  }
  int enqueue_frag_dots(const int *flag) {
    if (ts.nheavy == 0) return GB_OK;
    k_frag_dots<T, S><<<(ts.nheavy + FRAG_DOT_WARPS - 1) / FRAG_DOT_WARPS, FRAG_DOT_WARPS * 32, 0, ctx->stream>>>(ts, J, xs, flag);
    GB_LAUNCH(ctx);
    return GB_OK;
  }

  int ensure_camx() {
    if (camx_valid) return GB_OK;
    k_cam_precompute<T><<<(ts.Nc + 127) / 128, 128, 0, ctx->stream>>>(ts.Nc, cams, camx);
    GB_LAUNCH(ctx);
    camx_valid = true;
    return GB_OK;
  }
  // need_cost = false: the caller does not read chi2 of this linearisation (the accepted-step path of the LM loop
  // already has it from the trial step), so its cross-rank sum is skipped
  // imported.r != nullptr: the caller's own linearisation (gb_import_linearization) instead of a factor evaluation
  int enqueue_linearize(bool need_cost = true, ExtFactor imported = ExtFactor{nullptr, nullptr, nullptr, nullptr, nullptr}) {
    cudaStream_t st = ctx->stream;
    const bool multi = ctx->nranks > 1;
    if (prescaled) {
      GB_TRY(require(!multi && !imported.r && !ext_fn && !rb.Pu && rb.loss_kind == 0,
                     "bf16 Jacobian storage: single rank, built-in factor, default loss and identity precision only"));
    }
    if (imported.r) {
      k_linearize<T, S, true><<<ts.nst, TILE, smem_lin_bytes<T>(), st>>>(ts, camx, pts, obs, J, res, Cg, part18, cost_part, rb, imported);
    } else if (ext_fn) {
      GB_TRY(ensure_camx());
      GB_TRY(evaluate_external(true));
      k_linearize<T, S, true><<<ts.nst, TILE, smem_lin_bytes<T>(), st>>>(ts, camx, pts, obs, J, res, Cg, part18, cost_part, rb, ex);
    } else {
      GB_TRY(ensure_camx());
      k_linearize<T, S, false><<<ts.nst, TILE, smem_lin_bytes<T>(), st>>>(ts, camx, pts, obs, J, res, Cg, part18, cost_part, rb, ex);
    }
    GB_LAUNCH(ctx);
    GB_TRY(enqueue_frag_sum(9, Cg));
    // pre-scaled storage: this first pass only yields the rounded Jacobians and, from them, the true Jacobi scales
    const int son = prescaled ? (scale_on ? 1 : 0) : alg_scale();
    k_cam_reduce_lin<T><<<ts.Nc, 288, 0, st>>>(ts, part18, diagB, gc, multi ? 0 : 1, son, scale, b, fixed_c);
    GB_LAUNCH(ctx);
    k_sum_partials<<<1, 1024, 0, st>>>(cost_part, ts.ntiles, scalars, 0);
    GB_LAUNCH(ctx);
    if (multi) GB_TRY(exchange_push<T>(diagB, 2 * (size_t)dimc));
    // point scales and b_p (mu-independent part of k_point_prepare); W/h are refreshed by prepare.  Purely local: on
    // the multi-GPU path it runs while the peers' diag(B) | g_c are in flight
    k_point_prepare<T><<<(ts.Np + 255) / 256, 256, 0, st>>>(ts.Np, son, mu, use_identity, Cg, scale + dimc,
                                                           b + dimc, W, h, 1, fixed_p);
    GB_LAUNCH(ctx);
    if (multi) {
      GB_TRY(exchange_sum<T>(diagB, 2 * (size_t)dimc));
      if (need_cost) GB_TRY(exchange<double>(scalars, 1));
      k_cam_finish_lin<T><<<(int)((dimc + 255) / 256), 256, 0, st>>>((int)dimc, son, diagB, gc, scale, b, fixed_c);
      GB_LAUNCH(ctx);
    }
    if (prescaled) {
      // second pass (ops/linearize.hpp:140-180): J~ = (S)((T)J * s) in place, then the assembly from J~ with unit scales
      k_copy<T><<<4 * 148, 256, 0, st>>>((int64_t)dimH, scale, scale_true);
      GB_LAUNCH(ctx);
      ExtFactor rs{nullptr, nullptr, nullptr, nullptr, scale_true};
      k_linearize<T, S, false, true><<<ts.nst, TILE, smem_lin_bytes<T>(), st>>>(ts, camx, pts, obs, J, res, Cg, part18, cost_part, rb, rs);
      GB_LAUNCH(ctx);
      GB_TRY(enqueue_frag_sum(9, Cg));
      k_cam_reduce_lin<T><<<ts.Nc, 288, 0, st>>>(ts, part18, diagB, gc, 1, 0, scale, b);
      GB_LAUNCH(ctx);
      k_point_prepare<T><<<(ts.Np + 255) / 256, 256, 0, st>>>(ts.Np, 0, mu, use_identity, Cg, scale + dimc, b + dimc, W, h, 1);
      GB_LAUNCH(ctx);
    }
    GB_TRY(launch_check());
    linearized = true;
    prepared = solved = stepped = false;
    full_lin_valid = solved_full = false;
    return GB_OK;
  }

  // ---- full-system matrix-free PCG: PCGSolver + BlockJacobiPreconditioner (solver/pcg.hpp:61-232,
  //      preconditioner/block_jacobi.hpp:79-186).  The loop is driven from the host with blocking scalar reads,
  //      like the reference's; products and preconditioner run in the kernels of kernels.cuh. ----------------------
  int full_buffers() {
    if (full_alloc) return GB_OK;
    GB_TRY(require(ctx->nranks == 1, "the full-system PCG solver is single-rank"));
    const size_t n = (size_t)dimH;
    GB_TRY(dalloc(f_x, n)); GB_TRY(dalloc(f_r, n)); GB_TRY(dalloc(f_z, n));
    GB_TRY(dalloc(f_p, n)); GB_TRY(dalloc(f_v2, n)); GB_TRY(dalloc(f_xbak, n));
    GB_TRY(dalloc(f_upw, (size_t)WST<T>::value * ts.Np + 8));
    GB_TRY(dalloc(f_outp, 3 * (size_t)ts.Np));
    GB_TRY(dalloc(f_zero, (size_t)WST<T>::value * ts.Np + 8)); // W = 0, h = 0: k_prepare_cams then sums Jc^T Jc
    GB_TRY(dalloc(f_Bfull, (size_t)ts.Nc * 81)); GB_TRY(dalloc(f_MinvF, (size_t)ts.Nc * 81));
    full_blocks = (int)((dimH + 255) / 256);
    GB_TRY(dalloc(f_part, full_blocks)); GB_TRY(dalloc(f_rho, full_blocks));
    GB_TRY(dalloc(f_state, 1));
    GB_CUDA(ctx, cudaMallocHost((void **)&h_fstate, sizeof(FullState<T>)));
    full_alloc = true;
    return GB_OK;
  }
  // One batch of PCG iterations [k0, k1) enqueued without reading anything back: every kernel returns at once when the
  // device state says the solve has stopped.
  int enqueue_full_iterations(int64_t k0, int64_t k1, T tol, T ratio, int max_iter) {
    cudaStream_t st = ctx->stream;
    const int *flag = &f_state->done;
    for (int64_t k = k0; k < k1; k++) {
      // p = beta p + z (p = z at the start), u = D p ; v2 = J~^T J~ p + mu clamp(diag) p   (pcg.hpp:141-168)
      k_full_direction<T><<<full_blocks, 256, 0, st>>>(ts.Nc, ts.Np, k == 0 ? 1 : 0, f_state, f_p, f_z, scale, xs, f_upw);
      GB_LAUNCH(ctx);
      GB_TRY(launch_product<true>(f_upw, flag, f_outp));
      k_cam_reduce_spmv<T><<<ts.Nc, 288, 0, st>>>(ts, part9, scale, Ap_raw, 0, dterm, nullptr, Ap, dot_part, flag, pp, 0);
      GB_LAUNCH(ctx);
      k_full_finish_v2<T><<<full_blocks, 256, 0, st>>>(ts.Nc, ts.Np, mu, use_identity, Ap_raw, f_outp, f_p, scale, diagB, Cg,
                                                       f_v2, f_part, flag);
      GB_LAUNCH(ctx);
      k_full_scalar<T><<<1, 1024, 0, st>>>(2, f_part, full_blocks, f_state, tol, ratio, max_iter); // alpha = rz / p.v2
      GB_LAUNCH(ctx);
      k_full_update_xr<T><<<full_blocks, 256, 0, st>>>(dimH, f_state, f_p, f_v2, f_x, f_xbak, f_r, f_part);
      GB_LAUNCH(ctx);
      k_full_scalar<T><<<1, 1024, 0, st>>>(3, f_part, full_blocks, f_state, tol, ratio, max_iter); // 1 / |r|
      GB_LAUNCH(ctx);
      k_full_precond<T><<<full_blocks, 256, 0, st>>>(ts.Nc, ts.Np, f_state, f_MinvF, W, scale, f_r, f_z, f_part, 1);
      GB_LAUNCH(ctx);
      k_full_scalar<T><<<1, 1024, 0, st>>>(4, f_part, full_blocks, f_state, tol, ratio, max_iter); // r.z, tests, beta
      GB_LAUNCH(ctx);
    }
    return GB_OK;
  }
  // the state of the finished solve as the host sees it (after a synchronise of the stream)
  void full_finish_info() {
    if (!full_info_pending) return;
    full_info = gb_solve_info{};
    full_info.pcg_iterations = h_fstate->iter;
    full_info.rz_final = (double)h_fstate->rzn;
    full_info.stop_reason = h_fstate->reason;
    full_info_pending = false;
  }
  int solve_full(const gb_pcg_options *o) {
    cudaStream_t st = ctx->stream;
    GB_TRY(full_buffers());
    const size_t nbytes = (size_t)dimH * sizeof(T);
    if (!full_lin_valid) {
      // the 45 sums of Jc^T Jc per camera (mu-independent): the prepare pipeline with W = 0, h = 0
      k_prepare_cams<T, S><<<ts.nchunks, PC_THREADS, 0, st>>>(ts, J, f_zero, f_zero, part54);
      GB_LAUNCH(ctx);
      full_lin_valid = true;
      prepared = false; // part54 no longer holds the Schur sums
    }
    k_point_prepare<T><<<(ts.Np + 255) / 256, 256, 0, st>>>(ts.Np, alg_scale(), mu, use_identity, Cg, scale + dimc,
                                                           b + dimc, W, h, 0, fixed_p);
    GB_LAUNCH(ctx);
    k_full_cam_blocks<T><<<ts.Nc, 288, 0, st>>>(ts, part54, mu, use_identity, scale, f_Bfull, f_MinvF);
    GB_LAUNCH(ctx);
    const T tol = (T)o->tolerance, ratio = (T)o->rejection_ratio;
    const int max_iter = (int)o->max_iterations;
    // x = 0 ; r = b ; y = r / |r| ; z = M^-1 y ; rz = r.z   (pcg.hpp:93-121)
    GB_CUDA(ctx, cudaMemsetAsync(f_x, 0, nbytes, st));
    k_copy<T><<<4 * 148, 256, 0, st>>>((int64_t)dimH, b, f_r);
    GB_LAUNCH(ctx);
    k_vec_dot<T><<<full_blocks, 256, 0, st>>>(dimH, f_r, f_r, f_part);
    GB_LAUNCH(ctx);
    k_full_scalar<T><<<1, 1024, 0, st>>>(0, f_part, full_blocks, f_state, tol, ratio, max_iter);
    GB_LAUNCH(ctx);
    k_full_precond<T><<<full_blocks, 256, 0, st>>>(ts.Nc, ts.Np, f_state, f_MinvF, W, scale, f_r, f_z, f_part, 0);
    GB_LAUNCH(ctx);
    k_full_scalar<T><<<1, 1024, 0, st>>>(1, f_part, full_blocks, f_state, tol, ratio, max_iter);
    GB_LAUNCH(ctx);
    // Iterations in batches: the whole solve when it is short (the BAL protocol's 10 iterations: no read-back at all), else
    // 32 at a time with one look at the state in between, so that a converged solve does not pay for thousands of launches
    // that return at once.
    constexpr int64_t BATCH = 32;
    for (int64_t k0 = 0; k0 < max_iter; k0 += BATCH) {
      if (k0 > 0) {
        GB_TRY(store_host(h_fstate, f_state, sizeof(FullState<T>)));
        GB_CUDA(ctx, cudaStreamSynchronize(st));
        if (h_fstate->done) break;
      }
      GB_TRY(enqueue_full_iterations(k0, std::min<int64_t>(max_iter, k0 + BATCH), tol, ratio, max_iter));
    }
    k_full_restore<T><<<4 * 148, 256, 0, st>>>((int64_t)dimH, f_state, f_xbak, f_x);
    GB_LAUNCH(ctx);
    GB_TRY(store_host(h_fstate, f_state, sizeof(FullState<T>)));
    full_info_pending = true; // read by full_finish_info() after the caller's synchronise
    GB_TRY(launch_check());
    last_pcg = *o;
    solved_full = true;
    solved = false;
    stepped = false;
    return GB_OK;
  }
  // delta = x, rho partials, (apply) vertices += D x   (ops/update.hpp:9-31, levenberg_marquardt.hpp:34-41)
  int enqueue_step_full(bool apply) {
    cudaStream_t st = ctx->stream;
    k_cam_step<T><<<ncamblocks, 256, 0, st>>>((int)dimc, f_x, scale, b, mu, xs, cams, cams_bak, delta, rho_part + ts.ntiles,
                                              apply ? 1 : 0, apply_scale());
    GB_LAUNCH(ctx);
    if (apply) camx_valid = false;
    const int64_t n3 = 3 * (int64_t)ts.Np;
    const int nb = (int)((n3 + 255) / 256);
    k_full_point_step<T><<<nb, 256, 0, st>>>(n3, f_x + dimc, scale + dimc, b + dimc, mu, pts, pts_bak, delta + dimc, f_rho,
                                             apply ? 1 : 0, apply_scale() + dimc);
    GB_LAUNCH(ctx);
    k_sum_partials<<<1, 1024, 0, st>>>(f_rho, nb, scalars, 1);
    GB_LAUNCH(ctx);
    k_sum_partials<<<1, 1024, 0, st>>>(rho_part + ts.ntiles, ncamblocks, scalars, 2);
    GB_LAUNCH(ctx);
    GB_TRY(launch_check());
    return GB_OK;
  }

  // one launch of the Schur product kernel (exports, full-system solver, NCCL fallback, stage timers)
  template <bool FULL> int launch_product(const T *Wp, const int *flag, T *outp) {
    if (!FULL) GB_TRY(enqueue_frag_dots(flag)); // long tracks: their point sums come first
    k_schur_product2<T, S, FULL><<<ts.ncta, 2 * TILE, SchurSmem2<T, S>::TOTAL, ctx->stream>>>(ts, J, Wp, xs, part9, flag, outp);
    GB_LAUNCH(ctx);
    if (FULL) GB_TRY(enqueue_frag_sum(3, outp)); // ... or are completed afterwards
    return GB_OK;
  }
  int enqueue_prepare_tiles_only() {
    k_prepare_cams<T, S><<<ts.nchunks, PC_THREADS, 0, ctx->stream>>>(ts, J, W, h, part54);
    GB_LAUNCH(ctx);
    full_lin_valid = false; // part54 now holds the Schur sums, not the full-system solver's Jc^T Jc sums
    return GB_OK;
  }
  int enqueue_prepare() {
    cudaStream_t st = ctx->stream;
    k_point_prepare<T><<<(ts.Np + 255) / 256, 256, 0, st>>>(ts.Np, alg_scale(), mu, use_identity, Cg, scale + dimc,
                                                           b + dimc, W, h, 0, fixed_p);
    GB_LAUNCH(ctx);
    k_prepare_cams<T, S><<<ts.nchunks, PC_THREADS, 0, st>>>(ts, J, W, h, part54);
    GB_LAUNCH(ctx);
    full_lin_valid = false; // part54 now holds the Schur sums, not the full-system solver's Jc^T Jc sums
    const bool multi = ctx->nranks > 1;
    const int xchg = (multi && exchange_p2p<T>((size_t)ts.Nc * 54)) ? 1 : 0; // push / pull fused into the two launches
    k_cam_reduce_prepare<T><<<ts.Nc, 288, 0, st>>>(ts, part54, 0, sums54, multi ? 0 : 1, mu, use_identity, diagB, gc,
                                                   scale, Sdiag, Minv, bS, dterm, pp, xchg);
    GB_LAUNCH(ctx);
    if (multi) {
      if (!xchg) GB_TRY(allreduce_T(sums54, (size_t)ts.Nc * 54));
      k_cam_reduce_prepare<T><<<ts.Nc, 288, 0, st>>>(ts, part54, 1, sums54, 1, mu, use_identity, diagB, gc, scale, Sdiag,
                                                     Minv, bS, dterm, pp, xchg);
      GB_LAUNCH(ctx);
    }
    GB_TRY(launch_check());
    prepared = true;
    return GB_OK;
  }

  // Ap_raw = D (B - E W E^T) D v for the vector v whose scaled copy D v is in xs; with `pvec` also
  // Ap = Ap_raw + dterm pvec and the per-camera partials of pvec.Ap.  One product launch + the per-camera sum of its
  // partial rows: the form the exports (gb_schur_multiply), the stage timers and the NCCL fallback use.
  int enqueue_schur_product(const int *flag, const T *pvec) {
    cudaStream_t st = ctx->stream;
    const bool multi = ctx->nranks > 1;
    const int finish = (!multi && pvec) ? 1 : 0;
    GB_TRY(launch_product<false>(W, flag, nullptr));
    k_cam_reduce_spmv<T><<<ts.Nc, 288, 0, st>>>(ts, part9, scale, Ap_raw, finish, dterm, pvec, Ap, dot_part, flag, pp, 0);
    GB_LAUNCH(ctx);
    if (multi) GB_TRY(allreduce_T(Ap_raw, dimc)); // Ap and the dot partials are then formed inside k_pcg_update
    return GB_OK;
  }

  int ensure_explicit() {
    if (xs_ready) return GB_OK;
    hs.explicit_schur(xs_host);
    const HostStructure::ExplicitSchur &E = xs_host;
    xs_dev.Nc = ts.Nc;
    xs_dev.nblocks = (int32_t)E.rowidx.size();
    GB_TRY(upload(xs_dev.tptr, E.tptr));
    GB_TRY(upload(xs_dev.tup_a, E.tup_a)); GB_TRY(upload(xs_dev.tup_b, E.tup_b)); GB_TRY(upload(xs_dev.tup_p, E.tup_p));
    GB_TRY(upload(xs_dev.blk_row, E.blk_row)); GB_TRY(upload(xs_dev.blk_col, E.blk_col));
    GB_TRY(upload(xs_dev.diag_block, E.diag_block));
    GB_TRY(upload(xs_dev.row_ptr, E.row_ptr)); GB_TRY(upload(xs_dev.row_ent, E.row_ent)); GB_TRY(upload(xs_dev.row_other, E.row_other));
    GB_TRY(dalloc(xs_vals, (size_t)xs_dev.nblocks * 81));
    GB_TRY(dalloc(xs_p, (size_t)dimc));
    int per_sm = 0, sms = 0;
    GB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pcg_solve_explicit<T>, XS_THREADS, 0));
    GB_CUDA(ctx, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
    xs_grid = (int)std::min<int64_t>((int64_t)std::max(per_sm, 1) * sms, ts.Nc);
    GB_TRY(dalloc(xs_red, 2 * (size_t)xs_grid));
    GB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    xs_ready = true;
    return GB_OK;
  }
  // S at the current linearisation and damping, blocks in the order of gb_schur_structure (needs enqueue_prepare: W, Sdiag)
  int enqueue_schur_build() {
    GB_TRY(ensure_explicit());
    k_schur_build<T, S><<<(xs_dev.nblocks + 7) / 8, 256, 0, ctx->stream>>>(xs_dev, J, W, scale, Sdiag, xs_vals);
    GB_LAUNCH(ctx);
    return launch_check();
  }
  // Which form of the Schur complement a solve runs on.  GB_SCHUR_AUTO is the measured rule of DESIGN.md section 3
  // (scripts/schur_crossover.py on one B200, FP64; gpurun_out/r2o_crossover.log): per solve
  //     implicit  t = a_i + k (17.2 us + 43.9 us per million observations)
  //     explicit  t = a_i + 0.05 ms + 0.55 ns per (point, camera pair) tuple + k (9.8 us + 0.435 us per thousand blocks of S)
  // so the explicit form pays when the solve may run more than k* = build / saving-per-iteration PCG iterations: 11 / 24 /
  // 53 / 84 iterations measured at the Ladybug / Trafalgar / Dubrovnik / Venice shapes.  With the reference's BAL protocol
  // (10 iterations, examples/bal.cu:284-309) every BASELINE shape therefore runs matrix-free.  The number of blocks is
  // bounded by min(Nc (Nc + 1) / 2, tuples + Nc) without building the structure.
  double xs_kstar = -1.0;
  int choose_schur_mode(const gb_pcg_options *o) {
    if (o->schur_mode == GB_SCHUR_EXPLICIT || o->schur_mode == GB_SCHUR_IMPLICIT) return o->schur_mode;
    // Several ranks: every rank must run the SAME form (the two solve kernels exchange differently), and the rule below
    // looks at this rank's share only - on small shares ranks disagreed (8 ranks on a 7 k-observation problem: some chose
    // the stored S, the others waited for their exchange for ever).  The stored S does not shrink with the number of ranks
    // (points shard, camera pairs do not), so auto means matrix-free there; explicit stays available on request.
    if (ctx->nranks > 1) return GB_SCHUR_IMPLICIT;
    if (xs_kstar < 0.0) {
      double tuples = 0.0;
      for (int32_t p = 0; p < hs.Np; p++) {
        const double t = (double)(hs.pptr[p + 1] - hs.pptr[p]);
        tuples += 0.5 * t * (t - 1.0);
      }
      const double nc = (double)hs.Nc, blocks = std::min(0.5 * nc * (nc + 1.0), tuples + nc);
      const double scale_s = sizeof(S) == 8 ? 1.0 : 0.6; // FP32 Jacobians: the matrix-free pass is cheaper (measured 0.55-0.6)
      const double build_us = 50.0 + 0.55e-3 * tuples;
      const double saving_us = (17.2 + 43.9e-6 * (double)hs.M * scale_s) - (9.8 + 0.435e-3 * blocks);
      xs_kstar = saving_us > 0.5 ? build_us / saving_us : 1e30;
    }
    return (double)o->max_iterations > xs_kstar ? GB_SCHUR_EXPLICIT : GB_SCHUR_IMPLICIT;
  }

  // EigenSchurLDLTSolver::solve (eigen_schur.hpp:72-108) on the device: S (explicit blocks) -> dense -> Cholesky -> x_c
  int enqueue_direct(const gb_pcg_options *o) {
    cudaStream_t st = ctx->stream;
    GB_TRY(require(ctx->nranks == 1, "the direct Schur solver is single-rank"));
    tk_iters = 0;
    const int n = (int)dimc;
    if (dimc > 20000) return ctx->fail(GB_ERR_UNSUPPORTED, "direct Schur solve: 9 n_cams = %ld > 20000 (dense factorisation)", (long)dimc);
    GB_TRY(enqueue_schur_build());
    if (!dn_A) {
      GB_TRY(dalloc(dn_A, (size_t)n * n));
      GB_TRY(dalloc(dn_fail, 1));
      int per_sm = 0, sms = 0;
      GB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_cholesky<T>, CH_THREADS, 0));
      GB_CUDA(ctx, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
      dn_grid = std::max(1, std::min(per_sm, 2)) * sms;
    }
    GB_CUDA(ctx, cudaMemsetAsync(dn_A, 0, (size_t)n * n * sizeof(T), st));
    GB_CUDA(ctx, cudaMemsetAsync(dn_fail, 0, sizeof(int), st));
    k_dense_from_blocks<T><<<(xs_dev.nblocks + 7) / 8, 256, 0, st>>>(ts.Nc, xs_dev.nblocks, xs_dev.blk_row, xs_dev.blk_col, xs_vals, dn_A);
    GB_LAUNCH(ctx);
    {
      int nn = n;
      void *args[] = {(void *)&nn, (void *)&dn_A, (void *)&dn_fail};
      GB_CUDA(ctx, cudaLaunchCooperativeKernel((const void *)k_cholesky<T>, dim3(dn_grid), dim3(CH_THREADS), args, 0, st));
      GB_LAUNCH(ctx);
    }
    k_cholesky_solve<T><<<1, 1024, 0, st>>>(n, dn_A, bS, x, dn_fail);
    GB_LAUNCH(ctx);
    k_direct_state<T><<<1, 1, 0, st>>>(pcg_state, dn_fail);
    GB_LAUNCH(ctx);
    GB_TRY(store_host(h_state, pcg_state, sizeof(PcgState<T>)));
    GB_TRY(launch_check());
    last_pcg = *o;
    last_schur_mode = GB_SCHUR_EXPLICIT;
    solved = true;
    solved_full = false;
    stepped = false;
    return GB_OK;
  }

  int ensure_timing(int64_t max_iter) {
    const int need = (int)((max_iter + 1) * SOLVE_STAMPS);
    if (need > timing_cap) {
      GB_TRY(dalloc(d_timing, need));
      if (h_timing) cudaFreeHost(h_timing);
      GB_CUDA(ctx, cudaMallocHost((void **)&h_timing, need * sizeof(unsigned long long)));
      timing_cap = need;
    }
    return GB_OK;
  }

  int ensure_tk(int iters) {
    if (!tk_alpha) GB_TRY(dalloc(tk_alpha, TK_MAX));
    if (iters > tk_alloc) {
      GB_TRY(dalloc(tk_buf, (size_t)iters * 3 * ts.Np)); // (a smaller earlier buffer stays in `allocs` until the problem goes)
      tk_alloc = iters;
    }
    return GB_OK;
  }
  // PCGSchurSolver::solve (pcg_schur.hpp:79-168) after the Schur / preconditioner values: ONE cooperative launch
  // (k_pcg_solve).  Without peer memory between the ranks (GB_P2P=0, IPC unavailable) the iterations are separate
  // launches with an NCCL all-reduce of S p in between.
  int enqueue_pcg(const gb_pcg_options *o) {
    cudaStream_t st = ctx->stream;
    const T tol = (T)o->tolerance, ratio = (T)o->rejection_ratio;
    const int max_iter = (int)o->max_iterations;
    tk_iters = 0;
    last_schur_mode = choose_schur_mode(o);
    if (last_schur_mode == GB_SCHUR_EXPLICIT && (ctx->nranks == 1 || p2p_on)) {
      GB_TRY(enqueue_schur_build());
      const T *c_vals = xs_vals, *c_Minv = Minv, *c_bS = bS;
      PcgState<T> *stp = pcg_state;
      int multi = ctx->nranks > 1 ? 1 : 0;
      T tol_ = tol, ratio_ = ratio;
      int mi = max_iter;
      void *args[] = {(void *)&xs_dev, (void *)&c_vals, (void *)&c_Minv, (void *)&c_bS, (void *)&x, (void *)&xbak, (void *)&r,
                      (void *)&z, (void *)&xs_p, (void *)&Ap, (void *)&xs_red, (void *)&stp, (void *)&tol_, (void *)&ratio_,
                      (void *)&mi, (void *)&pp, (void *)&multi};
      GB_CUDA(ctx, cudaLaunchCooperativeKernel((const void *)k_pcg_solve_explicit<T>, dim3(xs_grid), dim3(XS_THREADS), args, 0, st));
      GB_LAUNCH(ctx);
      GB_TRY(store_host(h_state, pcg_state, sizeof(PcgState<T>)));
    } else if (ctx->nranks == 1 || p2p_on) {
      last_schur_mode = GB_SCHUR_IMPLICIT;
      if (profiling) GB_TRY(ensure_timing(o->max_iterations));
      // short solves (the BAL protocol: 10 iterations) keep their per-iteration point sums: the back-substitution then
      // reads 3 values per point and iteration instead of streaming the Jacobians once more
      ts.tk_cap = 0;
      if (max_iter >= 1 && max_iter <= TK_MAX) {
        GB_TRY(ensure_tk(max_iter));
        ts.tk = tk_buf; ts.tk_alpha = tk_alpha; ts.tk_cap = max_iter;
      }
      const S2 *c_J = J;
      const T *c_W = W, *c_scale = scale, *c_dterm = dterm, *c_Minv = Minv, *c_bS = bS;
      PcgState<T> *stp = pcg_state;
      int multi = ctx->nranks > 1 ? 1 : 0;
      unsigned long long *tim = profiling ? d_timing : nullptr;
      T tol_ = tol, ratio_ = ratio;
      int mi = max_iter;
      void *args[] = {(void *)&ts, (void *)&c_J, (void *)&c_W, (void *)&c_scale, (void *)&c_dterm, (void *)&c_Minv,
                      (void *)&c_bS, (void *)&x, (void *)&xbak, (void *)&r, (void *)&z, (void *)&pbuf, (void *)&part9,
                      (void *)&st_dot, (void *)&cta_red, (void *)&stp, (void *)&work_counter, (void *)&tol_, (void *)&ratio_,
                      (void *)&mi, (void *)&pp, (void *)&multi, (void *)&tim};
      GB_CUDA(ctx, cudaLaunchCooperativeKernel((const void *)k_pcg_solve<T, S>, dim3(solve_grid), dim3(SolveSmem<T, S>::THREADS), args,
                                               SolveSmem<T, S>::TOTAL, st));
      GB_LAUNCH(ctx);
      GB_TRY(store_host(h_state, pcg_state, sizeof(PcgState<T>)));
      if (profiling) GB_TRY(store_host(h_timing, d_timing, (size_t)(max_iter + 1) * SOLVE_STAMPS * sizeof(unsigned long long)));
      tk_iters = ts.tk_cap;
    } else {
      last_schur_mode = GB_SCHUR_IMPLICIT;
      GB_TRY(ensure_state_cap(o->max_iterations));
      const int gridc = (ts.Nc + PCG_CAMS - 1) / PCG_CAMS;
      k_pcg_init<T><<<gridc, 288, 0, st>>>(ts.Nc, bS, Minv, scale, x, r, z, pv, xs, rz_part);
      GB_LAUNCH(ctx);
      k_pcg_init_state<T><<<1, 1024, 0, st>>>(ts.Nc, rz_part, pcg_state, done_flag);
      GB_LAUNCH(ctx);
      const int nc = ts.Nc;
      // An iteration enqueued after the PCG has stopped returns at once but still costs its launches, so only as many
      // iterations as the previous solve needed, plus two, are enqueued at first; if that is fewer than max_iterations
      // the state is read back and the rest follows only when the PCG is still running (identical on all ranks).
      const int64_t first_batch = std::min<int64_t>(o->max_iterations, pcg_guess + 2);
      int64_t last = first_batch;
      for (int64_t k = 0; k < o->max_iterations; k++) {
        if (k == first_batch) {
          GB_TRY(store_host(h_state, pcg_state + 2 * k, sizeof(PcgState<T>)));
          GB_CUDA(ctx, cudaStreamSynchronize(st));
          if (h_state->done) break;
          last = o->max_iterations;
        }
        GB_TRY(enqueue_schur_product(done_flag, pv));
        PcgState<T> *stp = pcg_state + 2 * k;
        const T *c_Minv = Minv, *c_scale = scale, *c_dterm = dterm, *c_raw = Ap_raw;
        int pull = 0;
        T tol_ = tol, ratio_ = ratio;
        int mi = max_iter;
        void *args[] = {(void *)&nc, (void *)&stp, (void *)&tol_, (void *)&ratio_, (void *)&mi, (void *)&dot_part,
                        (void *)&Ap, (void *)&c_Minv, (void *)&c_scale, (void *)&x, (void *)&xbak, (void *)&r,
                        (void *)&z, (void *)&pv, (void *)&xs, (void *)&rz_part, (void *)&done_flag, (void *)&c_raw,
                        (void *)&c_dterm, (void *)&pp, (void *)&pull};
        GB_CUDA(ctx, cudaLaunchCooperativeKernel((const void *)k_pcg_update<T>, dim3(gridc), dim3(288), args, 0, st));
        GB_LAUNCH(ctx);
      }
      GB_TRY(store_host(h_state, pcg_state + 2 * last, sizeof(PcgState<T>)));
    }
    GB_TRY(launch_check());
    last_pcg = *o;
    solved = true;
    solved_full = false;
    stepped = false;
    return GB_OK;
  }

  // back-substitution (+ optional application of the step) ; fills delta and the rho partials
  // sums = false: the caller runs enqueue_cost(true) next, which sums the rho partials together with the cost
  int enqueue_step(bool apply, bool sums = true) {
    cudaStream_t st = ctx->stream;
    // camera part: delta_c = x, xs = D x, rho, (apply) cams += D x
    k_cam_step<T><<<ncamblocks, 256, 0, st>>>((int)dimc, x, scale, b, mu, xs, cams, cams_bak, delta, rho_part + ts.ntiles,
                                              apply ? 1 : 0, apply_scale());
    GB_LAUNCH(ctx);
    if (apply) camx_valid = false;
    if (tk_iters > 0) {
      // the solve kernel kept t_p(p_k) and alpha_k of its iterations: no pass over the Jacobians
      const int nb = (ts.Np + 255) / 256;
      k_backsubst_points<T><<<nb, 256, 0, st>>>(ts, tk_iters, W, h, scale + dimc, b + dimc, mu, pts, pts_bak, delta + dimc, rho_part,
                                               apply ? 1 : 0, apply_scale() + dimc);
      rho_pt_n = nb;
    } else {
      GB_TRY(enqueue_frag_dots(nullptr));
      k_backsubst_tiles<T, S><<<ts.ntiles, TILE, 0, st>>>(ts, J, W, xs, h, scale + dimc, b + dimc, mu, pts, pts_bak,
                                                          delta + dimc, rho_part, apply ? 1 : 0, apply_scale() + dimc);
      rho_pt_n = ts.ntiles;
    }
    GB_LAUNCH(ctx);
    if (sums) {
      k_sum_partials<<<1, 1024, 0, st>>>(rho_part, rho_pt_n, scalars, 1);
      GB_LAUNCH(ctx);
      k_sum_partials<<<1, 1024, 0, st>>>(rho_part + ts.ntiles, ncamblocks, scalars, 2);
      GB_LAUNCH(ctx);
    }
    GB_TRY(launch_check());
    return GB_OK;
  }

  int enqueue_cost(bool with_rho = false) {
    cudaStream_t st = ctx->stream;
    GB_TRY(ensure_camx());
    if (ext_fn) {
      GB_TRY(evaluate_external(false));
      k_cost_tiles<T, true><<<(ts.ntiles + COST_TILES - 1) / COST_TILES, TILE, 0, st>>>(ts, camx, pts, obs, cost_part, rb, ex);
    } else {
      k_cost_tiles<T, false><<<(ts.ntiles + COST_TILES - 1) / COST_TILES, TILE, 0, st>>>(ts, camx, pts, obs, cost_part, rb, ex);
    }
    GB_LAUNCH(ctx);
    if (with_rho && !solved_full) { // cost and both rho sums of the Schur path in one launch
      k_sum_partials3<<<3, 1024, 0, st>>>(SumJob{cost_part, ts.ntiles, 0}, SumJob{rho_part, rho_pt_n, 1},
                                          SumJob{rho_part + ts.ntiles, ncamblocks, 2}, scalars);
    } else {
      k_sum_partials<<<1, 1024, 0, st>>>(cost_part, ts.ntiles, scalars, 0);
    }
    GB_LAUNCH(ctx);
    GB_TRY(launch_check());
    return GB_OK;
  }

  int fetch_scalars() {
    if (ctx->nranks > 1) GB_TRY(exchange<double>(scalars, 2)); // cost and the point part of rho; the camera part is replicated
    GB_TRY(store_host(h_scalars, scalars, 3 * sizeof(double)));
    GB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return p2p_check();
  }

  int require(bool cond, const char *what) {
    if (!cond) return ctx->fail(GB_ERR_INVALID, "%s", what);
    return GB_OK;
  }

  // ---- public operations ---------------------------------------------------------------------------------
  int linearize(double *chi2) override {
    GB_TRY(require(have_obs && have_vertices, "gb_linearize needs observations and vertices"));
    scale_on = true; // (an imported linearisation, gb_import_linearization, runs with scales = 1)
    GB_TRY(enqueue_linearize());
    GB_TRY(store_host(h_scalars, scalars, sizeof(double)));
    GB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    GB_TRY(p2p_check());
    last_chi2 = (double)(T)h_scalars[0];
    if (chi2) *chi2 = last_chi2;
    return GB_OK;
  }
  int compute_cost(double *chi2) override {
    GB_TRY(require(have_obs && have_vertices, "gb_compute_cost needs observations and vertices"));
    GB_TRY(enqueue_cost());
    if (ctx->nranks > 1) GB_TRY(exchange<double>(scalars, 1));
    GB_TRY(store_host(h_scalars, scalars, sizeof(double)));
    GB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    GB_TRY(p2p_check());
    if (chi2) *chi2 = (double)(T)h_scalars[0];
    return GB_OK;
  }
  int d2h(void *dst, const void *src, size_t n) {
    GB_CUDA(ctx, cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToHost, ctx->stream));
    GB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return p2p_check();
  }
  int get_gradient(void *out) override {
    GB_TRY(require(linearized, "gb_get_gradient before gb_linearize"));
    return d2h(out, b, dimH * sizeof(T));
  }
  int get_scales(void *out) override {
    GB_TRY(require(linearized, "gb_get_scales before gb_linearize"));
    return d2h(out, apply_scale(), dimH * sizeof(T));
  }
  int get_residuals(void *out) override {
    GB_TRY(require(linearized, "gb_get_residuals before gb_linearize"));
    std::vector<T> tmp(2 * (size_t)hs.Mstore);
    GB_TRY(d2h(tmp.data(), res, tmp.size() * sizeof(T)));
    T *o = (T *)out;
    for (int64_t i = 0; i < hs.M; i++) {
      const int64_t u = hs.identity_perm ? i : hs.perm[i], sl = hs.slot_of_obs[i];
      o[2 * u] = tmp[2 * sl];
      o[2 * u + 1] = tmp[2 * sl + 1];
    }
    return GB_OK;
  }
  int get_jacobians(double *oc, double *op) override {
    GB_TRY(require(linearized, "gb_get_jacobians before gb_linearize"));
    std::vector<S> hj(2 * (size_t)NPLANES * hs.Mstore);
    GB_TRY(d2h(hj.data(), J, hj.size() * sizeof(S)));
    for (int64_t i = 0; i < hs.M; i++) {
      const int64_t u = hs.identity_perm ? i : hs.perm[i], sl = hs.slot_of_obs[i];
      const int64_t tile = sl / TILE, t = sl % TILE;
      for (int j = 0; j < 9; j++) {
        const int64_t e = ((tile * NPLANES + j) * TILE + t) * 2;
        oc[18 * u + 2 * j] = (double)hj[e];
        oc[18 * u + 2 * j + 1] = (double)hj[e + 1];
      }
      for (int j = 0; j < 3; j++) {
        const int64_t e = ((tile * NPLANES + 9 + j) * TILE + t) * 2;
        op[6 * u + 2 * j] = (double)hj[e];
        op[6 * u + 2 * j + 1] = (double)hj[e + 1];
      }
    }
    return GB_OK;
  }
  int hessian_values(void *out) override {
    GB_TRY(require(linearized, "gb_hessian_values before gb_linearize"));
    GB_TRY(require(ctx->nranks == 1, "gb_hessian_values is single-rank only"));
    if (prescaled) return ctx->fail(GB_ERR_UNSUPPORTED, "gb_hessian_values: not offered for bf16 storage");
    const int64_t nv = 81 * (int64_t)hs.Nc + 27 * hs.M + 9 * (int64_t)hs.Np;
    Scratch s_vals, s_bacc;
    GB_CUDA(ctx, cudaMalloc(&s_vals.p, nv * sizeof(S)));
    GB_CUDA(ctx, cudaMalloc(&s_bacc.p, 81 * (size_t)hs.Nc * sizeof(double)));
    S *vals = s_vals.as<S>();
    double *Bacc = s_bacc.as<double>();
    GB_CUDA(ctx, cudaMemsetAsync(Bacc, 0, 81 * (size_t)hs.Nc * sizeof(double), ctx->stream));
    const int64_t n = std::max<int64_t>(hs.M, hs.Np);
    k_hessian_export<T, S><<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ts, J, Cg, scale, scale + dimc, vals, Bacc);
    GB_LAUNCH(ctx);
    k_copy_B<S><<<(81 * hs.Nc + 255) / 256, 256, 0, ctx->stream>>>(81 * hs.Nc, Bacc, vals);
    GB_LAUNCH(ctx);
    return d2h(out, vals, nv * sizeof(S));
  }
  int set_damping(double m, int ident) override {
    mu = (T)m;
    use_identity = ident;
    prepared = solved = false;
    return GB_OK;
  }
  int solve(const gb_pcg_options *o, void *delta_host, gb_solve_info *info) override {
    GB_TRY(require(linearized, "gb_solve before gb_linearize"));
    GB_TRY(require(o && o->max_iterations >= 0 && o->max_iterations < (1 << 20), "bad PCG options"));
    GB_TRY(require(o->solver >= GB_SOLVER_PCG_SCHUR && o->solver <= GB_SOLVER_DIRECT_SCHUR, "unknown solver"));
    GB_TRY(require(o->schur_mode >= GB_SCHUR_AUTO && o->schur_mode <= GB_SCHUR_EXPLICIT, "unknown schur_mode"));
    if (o->solver == GB_SOLVER_PCG_FULL) {
      GB_TRY(solve_full(o));
      if (delta_host) {
        GB_TRY(enqueue_step_full(false));
        GB_TRY(d2h(delta_host, delta, dimH * sizeof(T)));
      }
      GB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      full_finish_info();
      if (info) *info = full_info;
      return GB_OK;
    }
    GB_TRY(enqueue_prepare());
    if (o->solver == GB_SOLVER_DIRECT_SCHUR) GB_TRY(enqueue_direct(o));
    else GB_TRY(enqueue_pcg(o));
    if (delta_host) {
      GB_TRY(enqueue_step(false));
      GB_TRY(d2h(delta_host, delta, dimH * sizeof(T)));
    }
    GB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    GB_TRY(p2p_check());
    pcg_guess = h_state->iter;
    if (info) {
      info->pcg_iterations = h_state->iter;
      info->rz_final = (double)h_state->rz;
      info->stop_reason = h_state->reason;
      info->schur_mode = last_schur_mode;
    }
    return GB_OK;
  }
  int solve_device(const gb_pcg_options *o, void *delta_dev, gb_solve_info *info) override {
    GB_TRY(require(linearized, "gb_solve_device before a linearisation"));
    GB_TRY(require(o && o->max_iterations >= 0 && o->max_iterations < (1 << 20), "bad PCG options"));
    GB_TRY(require(o->solver == GB_SOLVER_PCG_SCHUR, "gb_solve_device: the Schur PCG solver only"));
    GB_TRY(enqueue_prepare());
    GB_TRY(enqueue_pcg(o));
    GB_TRY(enqueue_step(false));
    if (delta_dev) {
      k_copy<T><<<4 * 148, 256, 0, ctx->stream>>>((int64_t)dimH, delta, (T *)delta_dev);
      GB_LAUNCH(ctx);
    }
    GB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    GB_TRY(p2p_check());
    pcg_guess = h_state->iter;
    if (info) {
      info->pcg_iterations = h_state->iter;
      info->rz_final = (double)h_state->rz;
      info->stop_reason = h_state->reason;
      info->schur_mode = last_schur_mode;
    }
    return GB_OK;
  }
  int get_schur_rhs(void *out) override {
    GB_TRY(require(linearized, "gb_get_schur_rhs before gb_linearize"));
    if (!prepared) GB_TRY(enqueue_prepare());
    return d2h(out, bS, dimc * sizeof(T));
  }
  int get_schur_diagonal(void *out) override {
    GB_TRY(require(linearized, "gb_get_schur_diagonal before gb_linearize"));
    if (!prepared) GB_TRY(enqueue_prepare());
    return d2h(out, Sdiag, (size_t)hs.Nc * 81 * sizeof(T));
  }
  int schur_multiply(const void *xin, void *yout) override {
    GB_TRY(require(linearized, "gb_schur_multiply before gb_linearize"));
    if (!prepared) GB_TRY(enqueue_prepare());
    // xs = D x (host side scaling keeps this export path free of extra kernels)
    std::vector<T> hscale(dimc), hx(hs.Nc * CAM_STRIDE, T(0)), hd(dimc), hy(dimc);
    GB_TRY(d2h(hscale.data(), scale, dimc * sizeof(T)));
    GB_TRY(d2h(hd.data(), dterm, dimc * sizeof(T)));
    const T *xi = (const T *)xin;
    for (int64_t c = 0; c < hs.Nc; c++)
      for (int k = 0; k < 9; k++) hx[c * CAM_STRIDE + k] = hscale[c * 9 + k] * xi[c * 9 + k];
    GB_CUDA(ctx, cudaMemcpyAsync(xs, hx.data(), hx.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    GB_TRY(enqueue_schur_product(nullptr, nullptr));
    GB_TRY(d2h(hy.data(), Ap_raw, dimc * sizeof(T)));
    T *yo = (T *)yout;
    for (int64_t i = 0; i < dimc; i++) yo[i] = hy[i] + hd[i] * xi[i];
    solved = false; // xs was overwritten
    return GB_OK;
  }
  // ---- explicit Schur complement: structure (host, cached) and values ------------------------------------------------
  std::vector<int64_t> s_colptr;
  std::vector<int32_t> s_rowidx;
  int schur_structure(int64_t *colptr, int64_t *rowidx, int64_t *nnz) override {
    if (s_colptr.empty()) hs.schur_structure(s_colptr, s_rowidx);
    if (nnz) *nnz = (int64_t)s_rowidx.size();
    if (colptr) std::copy(s_colptr.begin(), s_colptr.end(), colptr);
    if (rowidx) for (size_t i = 0; i < s_rowidx.size(); i++) rowidx[i] = s_rowidx[i];
    return GB_OK;
  }
  int schur_values(void *out) override {
    GB_TRY(require(linearized, "gb_schur_values before gb_linearize"));
    GB_TRY(require(ctx->nranks == 1, "gb_schur_values is single-rank only"));
    if (!prepared) GB_TRY(enqueue_prepare());
    GB_TRY(enqueue_schur_build());
    return d2h(out, xs_vals, (size_t)xs_dev.nblocks * 81 * sizeof(T));
  }
  // scalar upper CSC of S in the reference's layout (csc_utils.hpp:73-193): host-side conversion of the block values
  int schur_csc(int32_t *ptr, int32_t *idx, void *vals_out, int64_t *nnz_out) override {
    int64_t nb = 0;
    GB_TRY(schur_structure(nullptr, nullptr, &nb));
    const int64_t n = 9 * (int64_t)hs.Nc;
    // column 9 j + c holds, for every block (i, j) of block column j (rows ascending), rows 9 i + r with 9 i + r <= 9 j + c
    int64_t nnz = 0;
    for (int64_t j = 0; j < hs.Nc; j++) {
      const int64_t off = s_colptr[j + 1] - s_colptr[j] - 1; // blocks above the diagonal one
      nnz += 9 * (9 * off) + 45;
    }
    if (nnz_out) *nnz_out = nnz;
    if (nnz >= (int64_t(1) << 31)) return ctx->fail(GB_ERR_UNSUPPORTED, "scalar CSC of S needs 64-bit indices");
    std::vector<T> blocks;
    if (vals_out) {
      blocks.resize((size_t)nb * 81);
      GB_TRY(schur_values(blocks.data()));
    }
    if (!ptr && !idx && !vals_out) return GB_OK;
    T *vo = (T *)vals_out;
    int64_t w = 0;
    for (int64_t j = 0; j < hs.Nc; j++)
      for (int c = 0; c < 9; c++) {
        const int64_t col = 9 * j + c;
        if (ptr) ptr[col] = (int32_t)w;
        for (int64_t k = s_colptr[j]; k < s_colptr[j + 1]; k++) {
          const int64_t i = s_rowidx[k];
          for (int r = 0; r < 9 && 9 * i + r <= col; r++) {
            if (idx) idx[w] = (int32_t)(9 * i + r);
            if (vo) vo[w] = blocks[(size_t)k * 81 + r + 9 * c]; // column-major 9x9 block
            w++;
          }
        }
      }
    if (ptr) ptr[n] = (int32_t)w;
    return GB_OK;
  }
  int try_step(double *new_chi2, double *rho_den) override {
    GB_TRY(require(solved || solved_full, "gb_try_step before gb_solve"));
    if (solved_full) GB_TRY(enqueue_step_full(true));
    else GB_TRY(enqueue_step(true, false));
    GB_TRY(enqueue_cost(!solved_full));
    GB_TRY(fetch_scalars());
    stepped = true;
    if (new_chi2) *new_chi2 = (double)(T)h_scalars[0];
    if (rho_den) *rho_den = (double)((T)(h_scalars[1] + h_scalars[2]) + (T)1.0e-3);
    return GB_OK;
  }
  int enqueue_revert() {
    cudaStream_t st = ctx->stream;
    k_copy<T><<<32, 256, 0, st>>>((int64_t)ts.Nc * CAM_STRIDE, cams_bak, cams);
    GB_LAUNCH(ctx);
    camx_valid = false;
    k_copy<T><<<4 * 148, 256, 0, st>>>(3 * (int64_t)ts.Np, pts_bak, pts);
    GB_LAUNCH(ctx);
    return GB_OK;
  }
  int revert_step() override {
    GB_TRY(require(stepped, "gb_revert_step without a step"));
    GB_TRY(enqueue_revert());
    GB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    stepped = false;
    return p2p_check();
  }

  // optimizer::levenberg_marquardt (levenberg_marquardt.hpp:109-242)
  int lm(const gb_lm_options *o, gb_lm_result *res_out, double *traj) override {
    GB_TRY(require(have_obs && have_vertices, "gb_lm needs observations and vertices"));
    GB_TRY(require(o && o->iterations >= 0, "bad LM options"));
    GB_TRY(require(o->pcg.solver >= GB_SOLVER_PCG_SCHUR && o->pcg.solver <= GB_SOLVER_DIRECT_SCHUR, "unknown solver"));
    GB_TRY(require(o->pcg.max_iterations >= 0 && o->pcg.max_iterations < (1 << 20), "bad PCG options"));
    GB_TRY(require(o->pcg.schur_mode >= GB_SCHUR_AUTO && o->pcg.schur_mode <= GB_SCHUR_EXPLICIT, "unknown schur_mode"));
    const bool full = o->pcg.solver == GB_SOLVER_PCG_FULL, direct = o->pcg.solver == GB_SOLVER_DIRECT_SCHUR;
    if (full) GB_TRY(full_buffers());
    cudaStream_t st = ctx->stream;
    gb_lm_result R{};
    if (!scale_on) { scale_on = true; linearized = false; } // leave the imported-linearisation mode
    const bool resume = o->resume != 0 && linearized;
    T mu_l = (T)o->initial_damping, nu = o->initial_nu > 0 ? (T)o->initial_nu : T(2);
    use_identity = o->use_identity;
    profiling = o->profile_product != 0;
    double acc[5] = {0, 0, 0, 0, 0};
    float ms = 0.f;
    GB_CUDA(ctx, cudaEventRecord(ev[6], st));
    mu = mu_l;
    T chi2;
    if (!resume) {
      GB_CUDA(ctx, cudaEventRecord(ev[0], st));
      GB_TRY(enqueue_linearize());
      GB_CUDA(ctx, cudaEventRecord(ev[1], st));
      GB_TRY(store_host(h_scalars, scalars, sizeof(double)));
      GB_CUDA(ctx, cudaStreamSynchronize(st));
      cudaEventElapsedTime(&ms, ev[0], ev[1]);
      acc[0] += ms;
      chi2 = (T)h_scalars[0];
    } else {
      chi2 = (T)last_chi2;
    }
    R.initial_chi2 = (double)chi2;
    bool run = true;
    int num_bad = 0;          // levenberg_marquardt2: consecutive accepted steps with < 0.1 % decrease
    bool lin_pending = false; // an accepted step's re-linearisation is in flight; its time is read after the next sync
    int64_t it = 0;
    for (; it < o->iterations && run; it++) {
      mu = mu_l;
      GB_CUDA(ctx, cudaEventRecord(ev[0], st));
      if (!full) GB_TRY(enqueue_prepare());
      GB_CUDA(ctx, cudaEventRecord(ev[1], st));
      if (full) GB_TRY(solve_full(&o->pcg));
      else if (direct) GB_TRY(enqueue_direct(&o->pcg));
      else GB_TRY(enqueue_pcg(&o->pcg));
      GB_CUDA(ctx, cudaEventRecord(ev[2], st));
      if (full) GB_TRY(enqueue_step_full(true));
      else GB_TRY(enqueue_step(true, false));
      GB_CUDA(ctx, cudaEventRecord(ev[3], st));
      GB_TRY(enqueue_cost(!full));
      GB_CUDA(ctx, cudaEventRecord(ev[4], st));
      GB_TRY(fetch_scalars());
      if (lin_pending) { cudaEventElapsedTime(&ms, ev[8], ev[9]); acc[0] += ms; lin_pending = false; }
      cudaEventElapsedTime(&ms, ev[0], ev[1]); acc[1] += ms;
      cudaEventElapsedTime(&ms, ev[1], ev[2]); acc[2] += ms;
      cudaEventElapsedTime(&ms, ev[2], ev[3]); acc[3] += ms;
      cudaEventElapsedTime(&ms, ev[3], ev[4]); acc[4] += ms;
      T new_chi2 = (T)h_scalars[0];
      // PCGSchurSolver::solve always returns true (pcg_schur.hpp:167); the direct solver reports a failed factorisation
      // (eigen_schur.hpp:79-82), which the loop turns into a rejected step (levenberg_marquardt.hpp:181-187, :19-47)
      const bool solve_ok = !(direct && h_state->reason == 6);
      if (!solve_ok) new_chi2 = std::numeric_limits<T>::max();
      const T denom = solve_ok ? (T)(h_scalars[1] + h_scalars[2]) + (T)1.0e-3 : T(1);
      const T rho = (chi2 - new_chi2) / denom;
      if (full) full_finish_info(); // (the stream was synchronised by fetch_scalars)
      const int64_t k_exec = full ? full_info.pcg_iterations : h_state->iter;
      if (!full) pcg_guess = k_exec;
      R.pcg_iterations_total += k_exec;
      if (profiling && !full && h_timing && (ctx->nranks == 1 || p2p_on)) {
        // phase stamps of CTA 0 (globaltimer, ns): product phase = P start .. after barrier A (every CTA has finished
        // its super-tiles), rest of the iteration = after barrier A .. after barrier B
        for (int64_t k = 0; k < k_exec; k++) {
          const unsigned long long *tk = h_timing + k * SOLVE_STAMPS;
          if (tk[5] <= tk[0]) continue; // an iteration that stopped before barrier B (bad denominator)
          R.product_seconds += 1e-9 * (double)(tk[2] - tk[0]);
          R.update_seconds += 1e-9 * (double)(tk[5] - tk[2]);
          R.product_launches++;
          for (int ph = 0; ph < 5; ph++) R.pcg_phase_seconds[ph] += 1e-9 * (double)(tk[ph + 1] - tk[ph]);
          if (tk[6] > tk[2]) R.pcg_phase_seconds[5] += 1e-9 * (double)(tk[6] - tk[2]); // multi-GPU: own sums pushed
        }
      }
      const bool accepted_now = solve_ok && std::isfinite((double)new_chi2) && rho > T(0);
      if (accepted_now) {
        double alpha = 1.0 - std::pow(2.0 * (double)rho - 1.0, 3);
        alpha = std::max(std::min(alpha, 2.0 / 3.0), 1.0 / 3.0);
        mu_l *= (T)alpha;
        nu = T(2);
        mu = mu_l;
        if (o->defer_final_linearize && it + 1 >= o->iterations) {
          // last iteration of this call: whoever needs the linearisation at the new point next computes it
          linearized = prepared = solved = solved_full = false;
        } else {
          // no host synchronisation here: the next iteration is enqueued while this linearisation runs
          GB_CUDA(ctx, cudaEventRecord(ev[8], st));
          GB_TRY(enqueue_linearize(false));
          GB_CUDA(ctx, cudaEventRecord(ev[9], st));
          lin_pending = true;
        }
        R.accepted++;
      } else {
        GB_TRY(enqueue_revert());
        mu_l *= nu;
        nu *= T(2);
        new_chi2 = chi2;
        R.rejected++;
        solved = solved_full = false;
      }
      if (traj) {
        traj[4 * it + 0] = (double)chi2; traj[4 * it + 1] = (double)new_chi2;
        traj[4 * it + 2] = (double)mu_l; traj[4 * it + 3] = (double)k_exec;
      }
      if (o->verbose)
        printf("%6ld %22.12g %22.12g %16.8g  pcg %ld\n", (long)it, (double)chi2, (double)new_chi2, (double)mu_l, (long)k_exec);
      const T initial_chi2 = chi2;
      chi2 = new_chi2;
      if (!std::isfinite((double)mu_l)) { run = false; R.termination = GB_LM_DAMPING_NOT_FINITE; }
      if (rho == T(0)) { it++; R.termination = GB_LM_RHO_ZERO; break; }
      if (o->stop_flag && *o->stop_flag) { it++; R.termination = GB_LM_STOP_FLAG; break; }
      if (o->early_stop && accepted_now) { // levenberg_marquardt.hpp:403-413
        if ((initial_chi2 - new_chi2) * T(1.0e3) < initial_chi2) num_bad++;
        else num_bad = 0;
        if (num_bad >= 3) { it++; R.termination = GB_LM_EARLY_STOP; break; }
      }
    }
    GB_CUDA(ctx, cudaEventRecord(ev[7], st));
    GB_CUDA(ctx, cudaStreamSynchronize(st));
    GB_TRY(p2p_check());
    if (lin_pending) { cudaEventElapsedTime(&ms, ev[8], ev[9]); acc[0] += ms; lin_pending = false; }
    cudaEventElapsedTime(&ms, ev[6], ev[7]);
    stepped = false;
    profiling = false;
    last_chi2 = (double)chi2;
    R.iterations = it;
    R.final_chi2 = (double)chi2;
    R.final_damping = (double)mu_l;
    R.final_nu = (double)nu;
    R.seconds_total = ms * 1e-3;
    R.seconds_linearize = acc[0] * 1e-3; R.seconds_prepare = acc[1] * 1e-3; R.seconds_pcg = acc[2] * 1e-3;
    R.seconds_backsubst = acc[3] * 1e-3; R.seconds_cost = acc[4] * 1e-3;
    if (res_out) *res_out = R;
    return GB_OK;
  }

  int time_stage(int stage, int reps, double *ms_avg) override {
    GB_TRY(require(have_obs && have_vertices, "gb_time_stage needs observations and vertices"));
    GB_TRY(require(reps > 0, "repetitions must be positive"));
    cudaStream_t st = ctx->stream;
    if (!linearized) GB_TRY(enqueue_linearize());
    if (stage >= 2 && !prepared) GB_TRY(enqueue_prepare());
    if (stage == 2 || stage == 3 || stage == 5 || stage == 6) {
      tk_iters = 0; // x comes from the initialisation below, not from a solve: the Jacobian-streaming back-substitution
      // a defined vector in xs / x: one PCG initialisation
      const int gridc = (ts.Nc + PCG_CAMS - 1) / PCG_CAMS;
      k_pcg_init<T><<<gridc, 288, 0, st>>>(ts.Nc, bS, Minv, scale, x, r, z, pv, xs, rz_part);
      GB_LAUNCH(ctx);
    }
    GB_CUDA(ctx, cudaStreamSynchronize(st));
    GB_CUDA(ctx, cudaEventRecord(ev[0], st));
    for (int i = 0; i < reps; i++) {
      switch (stage) {
      case 0: GB_TRY(enqueue_linearize()); break;
      case 1: GB_TRY(enqueue_prepare()); break;
      case 2: GB_TRY(enqueue_schur_product(nullptr, pv)); break;
      case 3: GB_TRY(enqueue_step(false)); break;
      case 4: GB_TRY(enqueue_cost()); break;
      case 5: // the product kernel alone
        GB_TRY(launch_product<false>(W, nullptr, nullptr));
        break;
      case 6: // the per-camera reduction of its partial rows alone (single-GPU form)
        k_cam_reduce_spmv<T><<<ts.Nc, 288, 0, st>>>(ts, part9, scale, Ap_raw, 1, dterm, pv, Ap, dot_part, nullptr, pp, 0);
        GB_LAUNCH(ctx);
        break;
      case 7: GB_TRY(enqueue_prepare_tiles_only()); break;
      default: return ctx->fail(GB_ERR_INVALID, "unknown stage %d", stage);
      }
    }
    GB_CUDA(ctx, cudaEventRecord(ev[1], st));
    GB_CUDA(ctx, cudaStreamSynchronize(st));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ev[0], ev[1]);
    *ms_avg = (double)ms / reps;
    if (stage == 0) { prepared = solved = false; }
    solved = false;
    return GB_OK;
  }
};

} // namespace gb

struct gb_problem {
  gb::ProblemBase *impl;
};

// ---------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------
extern "C" {

int gb_version(void) { return GB_VERSION; }

int gb_context_create(int device, gb_context **out) {
  if (!out) return GB_ERR_INVALID;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) return GB_ERR_NO_DEVICE;
  if (device < 0 || device >= n) return GB_ERR_INVALID;
  if (cudaSetDevice(device) != cudaSuccess) return GB_ERR_CUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return GB_ERR_CUDA;
  if (prop.major != 10) return GB_ERR_UNSUPPORTED; // kernels are built for sm_100a only
  gb_context *c = new gb_context();
  c->device = device;
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete c;
    return GB_ERR_CUDA;
  }
  *out = c;
  return GB_OK;
}

int gb_context_create_on_stream(int device, void *cuda_stream, gb_context **out) {
  const int rc = gb_context_create(device, out);
  if (rc != GB_OK) return rc;
  cudaStreamDestroy((*out)->stream);
  (*out)->stream = (cudaStream_t)cuda_stream;
  (*out)->owns_stream = false;
  return GB_OK;
}

int gb_context_destroy(gb_context *ctx) {
  if (!ctx) return GB_OK;
  cudaSetDevice(ctx->device);
  if (ctx->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(ctx->comm);
  if (ctx->stream && ctx->owns_stream) cudaStreamDestroy(ctx->stream);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  delete ctx;
  return GB_OK;
}

const char *gb_last_error(const gb_context *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int gb_comm_unique_id(void *id128) {
  if (!id128) return GB_ERR_INVALID;
  if (!g_nccl.load()) return GB_ERR_NCCL;
  ncclUniqueId_gb id;
  if (g_nccl.GetUniqueId(&id) != 0) return GB_ERR_NCCL;
  memcpy(id128, &id, 128);
  return GB_OK;
}

int gb_comm_init(gb_context *ctx, int nranks, int rank, const void *id128) {
  if (!ctx || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return GB_ERR_INVALID;
  if (!g_nccl.load()) return ctx->fail(GB_ERR_NCCL, "libnccl.so.2 not found");
  GB_CUDA(ctx, cudaSetDevice(ctx->device));
  ncclUniqueId_gb id;
  memcpy(&id, id128, 128);
  const int rc = g_nccl.CommInitRank(&ctx->comm, nranks, id, rank);
  if (rc != 0) return ctx->fail(GB_ERR_NCCL, "ncclCommInitRank failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
  ctx->nranks = nranks;
  ctx->rank = rank;
  return GB_OK;
}

int gb_problem_create(gb_context *ctx, const gb_problem_desc *d, gb_problem **out) {
  if (!ctx) return GB_ERR_INVALID;
  if (!d || !out || !d->camera_index || !d->point_index) return ctx->fail(GB_ERR_INVALID, "null argument");
  *out = nullptr;
  gb::ProblemBase *impl = nullptr;
  if (d->precision_T == GB_F64 && d->precision_S == GB_F64) impl = new gb::Problem<double, double>();
  else if (d->precision_T == GB_F32 && d->precision_S == GB_F32) impl = new gb::Problem<float, float>();
  else if (d->precision_T == GB_F64 && d->precision_S == GB_F32) impl = new gb::Problem<double, float>();
  else if (d->precision_T == GB_F64 && d->precision_S == GB_BF16) impl = new gb::Problem<double, __nv_bfloat16>();
  else return ctx->fail(GB_ERR_UNSUPPORTED, "precision (T=%d,S=%d) not supported", d->precision_T, d->precision_S);
  impl->ctx = ctx;
  const std::string why = impl->hs.build(d->num_cameras, d->num_points, d->num_observations, d->camera_index,
                                         d->point_index, d->tile_size, d->slot_cap, d->super_tile_observations,
                                         (d->flags & GB_FLAG_PARTITION) != 0, (d->flags & GB_FLAG_HOST_TABLES) == 0);
  if (!why.empty()) {
    delete impl;
    return ctx->fail(GB_ERR_UNSUPPORTED, "structure: %s", why.c_str());
  }
  const int rc = impl->init();
  if (rc != GB_OK) {
    delete impl;
    return rc;
  }
  *out = new gb_problem{impl};
  return GB_OK;
}

int gb_problem_destroy(gb_problem *p) {
  if (p) {
    delete p->impl;
    delete p;
  }
  return GB_OK;
}

struct gb_structure {
  gb::HostStructure hs;
};

static void fill_info(const gb::HostStructure &h, int64_t info[12], int64_t bytes) {
  info[0] = h.ntiles();
  info[1] = h.nrows();
  info[2] = h.max_track;
  info[3] = 9 * (int64_t)h.Nc + 3 * (int64_t)h.Np;
  info[4] = (int64_t)h.Nc + h.M + h.Np;
  info[5] = 81 * (int64_t)h.Nc + 27 * h.M + 9 * (int64_t)h.Np;
  info[6] = bytes;
  info[7] = h.M;
  info[8] = h.nst();
  info[9] = h.nseg_total;
  info[10] = h.Mstore;
  info[11] = 0;
}

int gb_structure_create(const gb_problem_desc *d, gb_structure **out, char *errbuf, int errlen) {
  if (!d || !out || !d->camera_index || !d->point_index) return GB_ERR_INVALID;
  *out = nullptr;
  gb_structure *s = new gb_structure();
  const std::string why = s->hs.build(d->num_cameras, d->num_points, d->num_observations, d->camera_index,
                                      d->point_index, d->tile_size, d->slot_cap, d->super_tile_observations,
                                      (d->flags & GB_FLAG_PARTITION) != 0);
  if (!why.empty()) {
    if (errbuf && errlen > 0) snprintf(errbuf, errlen, "%s", why.c_str());
    delete s;
    return GB_ERR_UNSUPPORTED;
  }
  *out = s;
  return GB_OK;
}
int gb_structure_destroy(gb_structure *s) { delete s; return GB_OK; }
int gb_structure_info(const gb_structure *s, int64_t info[12]) {
  if (!s || !info) return GB_ERR_INVALID;
  fill_info(s->hs, info, 0);
  return GB_OK;
}
int gb_structure_array(const gb_structure *s, int which, void *out, int64_t *count) {
  if (!s || !count) return GB_ERR_INVALID;
  if (which == 14 || which == 15) const_cast<gb_structure *>(s)->hs.materialize_tables();
  const gb::HostStructure &h = s->hs;
  const std::vector<int32_t> *v32[] = {&h.cam_idx, &h.pt_idx, &h.pptr, &h.tile_obs, &h.tile_pt, &h.st_tile,
                                       &h.st_row, &h.row_cam, &h.cam_row_ptr, &h.cam_row_list, &h.slot_of_obs};
  if (which >= 0 && which < 11) {
    *count = (int64_t)v32[which]->size();
    if (out) memcpy(out, v32[which]->data(), v32[which]->size() * sizeof(int32_t));
  } else if (which == 11) {
    *count = (int64_t)h.rank.size();
    if (out) memcpy(out, h.rank.data(), h.rank.size());
  } else if (which == 12) {
    *count = (int64_t)h.perm.size();
    if (out) memcpy(out, h.perm.data(), h.perm.size() * sizeof(int64_t));
  } else if (which == 13) {
    *count = (int64_t)h.ometa.size();
    if (out) memcpy(out, h.ometa.data(), h.ometa.size() * sizeof(uint32_t));
  } else if (which == 14) {
    *count = (int64_t)h.seg_tab.size();
    if (out) memcpy(out, h.seg_tab.data(), h.seg_tab.size() * sizeof(uint32_t));
  } else if (which == 15) {
    *count = (int64_t)h.pt_tab.size();
    if (out) memcpy(out, h.pt_tab.data(), h.pt_tab.size() * sizeof(uint16_t));
  } else if (which == 16) {
    *count = (int64_t)h.tmeta.size() * 8;
    if (out) memcpy(out, h.tmeta.data(), h.tmeta.size() * sizeof(gb::TileMeta));
  } else if (which == 17) {
    *count = (int64_t)h.trec.size();
    if (out) memcpy(out, h.trec.data(), h.trec.size());
  } else if (which >= 18 && which <= 20) {
    const std::vector<int32_t> &v = which == 18 ? h.tile_cam : (which == 19 ? h.cm_slot : h.cm_pt);
    *count = (int64_t)v.size();
    if (out) memcpy(out, v.data(), v.size() * sizeof(int32_t));
  } else if (which >= 21 && which <= 23) { // long tracks: fragment tiles, their points, point -> fragment range
    const std::vector<int32_t> &v = which == 21 ? h.frag_tile : (which == 22 ? h.hv_pt : h.hv_ptr);
    *count = (int64_t)v.size();
    if (out) memcpy(out, v.data(), v.size() * sizeof(int32_t));
  } else {
    return GB_ERR_INVALID;
  }
  return GB_OK;
}
int gb_structure_schur(const gb_structure *s, int64_t *cp, int64_t *ri, int64_t *nnz) {
  if (!s) return GB_ERR_INVALID;
  std::vector<int64_t> colptr;
  std::vector<int32_t> rowidx;
  s->hs.schur_structure(colptr, rowidx);
  if (nnz) *nnz = (int64_t)rowidx.size();
  if (cp) std::copy(colptr.begin(), colptr.end(), cp);
  if (ri) for (size_t i = 0; i < rowidx.size(); i++) ri[i] = rowidx[i];
  return GB_OK;
}
int gb_structure_hessian(const gb_structure *s, int64_t *cp, int64_t *ri, int64_t *off) {
  if (!s || !cp || !ri || !off) return GB_ERR_INVALID;
  s->hs.hessian_structure(cp, ri, off);
  return GB_OK;
}

int gb_problem_info(const gb_problem *p, int64_t info[12]) {
  if (!p || !info) return GB_ERR_INVALID;
  fill_info(p->impl->hs, info, p->impl->device_bytes());
  info[11] = p->impl->exchange_mode();
  return GB_OK;
}

#define GB_P(p) \
  if (!(p) || !(p)->impl) return GB_ERR_INVALID; \
  cudaSetDevice((p)->impl->ctx->device)

int gb_set_observations(gb_problem *p, const void *o) { GB_P(p); if (!o) return GB_ERR_INVALID; return p->impl->set_observations(o); }
int gb_stage_observations_async(gb_problem *p, const void *o, int slot) { GB_P(p); if (!o) return GB_ERR_INVALID; return p->impl->stage_observations_async(o, slot); }
int gb_commit_observations(gb_problem *p, int slot) { GB_P(p); return p->impl->commit_observations(slot); }
int gb_set_vertices(gb_problem *p, const void *c, const void *q) { GB_P(p); if (!c || !q) return GB_ERR_INVALID; return p->impl->set_vertices(c, q); }
int gb_get_vertices(gb_problem *p, void *c, void *q) { GB_P(p); return p->impl->get_vertices(c, q); }
int gb_set_observations_device(gb_problem *p, const void *o) { GB_P(p); if (!o) return GB_ERR_INVALID; return p->impl->set_observations_device(o); }
int gb_set_vertices_device(gb_problem *p, const void *c, const void *q) { GB_P(p); if (!c || !q) return GB_ERR_INVALID; return p->impl->set_vertices_device(c, q); }
int gb_get_vertices_device(gb_problem *p, void *c, void *q) { GB_P(p); return p->impl->get_vertices_device(c, q); }
int gb_import_linearization(gb_problem *p, const void *r, const void *jc, const void *jp) {
  GB_P(p);
  if (!r || !jc || !jp) return GB_ERR_INVALID;
  return p->impl->import_linearization(r, jc, jp);
}
int gb_solve_device(gb_problem *p, const gb_pcg_options *o, void *d, gb_solve_info *i) { GB_P(p); return p->impl->solve_device(o, d, i); }
int gb_set_factor(gb_problem *p, gb_factor_fn fn, void *user) { GB_P(p); return p->impl->set_factor(fn, user); }
int gb_set_loss(gb_problem *p, int kind, double delta) { GB_P(p); return p->impl->set_loss(kind, delta); }
int gb_set_precision(gb_problem *p, const void *P) { GB_P(p); return p->impl->set_precision(P); }
int gb_problem_structure_array(gb_problem *p, int which, void *out, int64_t *count) {
  GB_P(p);
  if (!count) return GB_ERR_INVALID;
  return p->impl->structure_array(which, out, count);
}
int gb_set_fixed(gb_problem *p, const uint8_t *fc, const uint8_t *fp) { GB_P(p); return p->impl->set_fixed(fc, fp); }
int gb_hessian_structure(const gb_problem *p, int64_t *cp, int64_t *ri, int64_t *off) {
  if (!p || !cp || !ri || !off) return GB_ERR_INVALID;
  p->impl->hs.hessian_structure(cp, ri, off);
  return GB_OK;
}
int gb_linearize(gb_problem *p, double *chi2) { GB_P(p); return p->impl->linearize(chi2); }
int gb_compute_cost(gb_problem *p, double *chi2) { GB_P(p); return p->impl->compute_cost(chi2); }
int gb_get_gradient(gb_problem *p, void *b) { GB_P(p); return p->impl->get_gradient(b); }
int gb_get_scales(gb_problem *p, void *s) { GB_P(p); return p->impl->get_scales(s); }
int gb_get_residuals(gb_problem *p, void *r) { GB_P(p); return p->impl->get_residuals(r); }
int gb_get_jacobians(gb_problem *p, double *a, double *b) { GB_P(p); return p->impl->get_jacobians(a, b); }
int gb_hessian_values(gb_problem *p, void *v) { GB_P(p); return p->impl->hessian_values(v); }
int gb_set_damping(gb_problem *p, double mu, int id) { GB_P(p); return p->impl->set_damping(mu, id); }
int gb_solve(gb_problem *p, const gb_pcg_options *o, void *d, gb_solve_info *i) { GB_P(p); return p->impl->solve(o, d, i); }
int gb_get_schur_rhs(gb_problem *p, void *b) { GB_P(p); return p->impl->get_schur_rhs(b); }
int gb_get_schur_diagonal(gb_problem *p, void *b) { GB_P(p); return p->impl->get_schur_diagonal(b); }
int gb_schur_multiply(gb_problem *p, const void *x, void *y) { GB_P(p); return p->impl->schur_multiply(x, y); }
int gb_schur_structure(gb_problem *p, int64_t *cp, int64_t *ri, int64_t *nnz) { GB_P(p); return p->impl->schur_structure(cp, ri, nnz); }
int gb_schur_values(gb_problem *p, void *v) { GB_P(p); if (!v) return GB_ERR_INVALID; return p->impl->schur_values(v); }
int gb_schur_csc(gb_problem *p, int32_t *ptr, int32_t *idx, void *v, int64_t *nnz) { GB_P(p); return p->impl->schur_csc(ptr, idx, v, nnz); }
int gb_try_step(gb_problem *p, double *c, double *r) { GB_P(p); return p->impl->try_step(c, r); }
int gb_revert_step(gb_problem *p) { GB_P(p); return p->impl->revert_step(); }
int gb_lm(gb_problem *p, const gb_lm_options *o, gb_lm_result *r, double *t) { GB_P(p); return p->impl->lm(o, r, t); }
int64_t gb_kernel_launches(const gb_context *ctx) { return ctx ? ctx->launches : 0; }
int gb_time_stage(gb_problem *p, int stage, int reps, double *ms) { GB_P(p); if (!ms) return GB_ERR_INVALID; return p->impl->time_stage(stage, reps, ms); }

} // extern "C"

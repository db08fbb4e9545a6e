/* graphite_b200.h — C ABI of the B200-native Levenberg-Marquardt inner loop for bundle adjustment.
 *
 * This is the drop-in boundary for ONE hot path of sfu-rsl/graphite: the LM loop with the
 * Schur-complement PCG solver on BAL-type problems (camera 9-dof / point 3-dof / 2-d reprojection
 * factors).  Every entry point names the reference interface it replaces (paths relative to the
 * reference tree).  Signatures carry plain pointers and sizes only; there are no C++ or torch types.
 *
 * Conventions
 *   - every function returns 0 on success or a negative gb_status; the message for the last failure
 *     on a context is gb_last_error(ctx).  Nothing throws.  There is NO CPU fallback: without a
 *     CUDA device (or when built without a matching kernel image) gb_context_create fails.
 *   - "host" pointers are caller-owned host memory (pinned or pageable); copies are issued on the
 *     context stream and completed before the call returns unless the name ends in _async.
 *   - T = graph precision (vertices, residuals, gradient, step, PCG vectors); S = linear-system
 *     precision (stored Jacobians).  (include/graphite/graph.hpp:25-29)
 *   - block order: cameras (vertex id = camera index) then points (vertex id = n_cams + point index),
 *     i.e. non-eliminated vertices by ascending id, then eliminated ones (graph.hpp:112-147).
 *   - the step delta_x is in the Jacobi-scaled space, length 9*n_cams + 3*n_points, cameras first
 *     (Solver::solve contract, solver/solver.hpp:24; pcg_schur.hpp:79-168).
 */
#ifndef GRAPHITE_B200_H
#define GRAPHITE_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GB_VERSION 100

typedef enum {
  GB_OK = 0,
  GB_ERR_INVALID = -1,     /* bad argument / call order */
  GB_ERR_CUDA = -2,        /* CUDA runtime failure (message has the CUDA error string) */
  GB_ERR_UNSUPPORTED = -3, /* precision combination or problem shape not supported */
  GB_ERR_NCCL = -4,
  GB_ERR_NO_DEVICE = -5
} gb_status;

/* GB_BF16: linear-system precision S only (stored Jacobians), with T = GB_F64 - the reference's low-precision mode
 * (types.hpp:10-19, examples/bal.cu:186-236, `--precision FP64-BF16`): Jacobians are rounded to bf16 when they are
 * evaluated and again after Jacobi scaling (ops/linearize.hpp:43-64, 140-180), every product accumulates in T. */
typedef enum { GB_F32 = 0, GB_F64 = 1, GB_BF16 = 2 } gb_dtype;

typedef struct gb_context gb_context;
typedef struct gb_problem gb_problem;

/* Replaces: cudaSetDevice + StreamPool (examples/bal.cu:53,251; include/graphite/stream.hpp). */
int gb_context_create(int device, gb_context **out);
/* The same on a stream the caller owns (cudaStream_t passed as void*; the reference hands its StreamPool to every Solver
 * call, solver/solver.hpp:16-24): all work of the context's problems is enqueued on it.  The stream must outlive the context. */
int gb_context_create_on_stream(int device, void *cuda_stream, gb_context **out);
int gb_context_destroy(gb_context *ctx);
const char *gb_last_error(const gb_context *ctx);
int gb_version(void);

/* Multi-GPU (new; the reference is single-GPU).  One process per GPU.  Rank 0 calls gb_comm_unique_id,
 * the 128 bytes are broadcast by the host program (torch.distributed / MPI), then every rank calls
 * gb_comm_init.  Points are partitioned across ranks by the caller; cameras are replicated. */
int gb_comm_unique_id(void *id128);
int gb_comm_init(gb_context *ctx, int nranks, int rank, const void *id128);

typedef struct {
  int32_t precision_T;       /* gb_dtype */
  int32_t precision_S;       /* gb_dtype; (F64,F64), (F32,F32), (F64,F32), (F64,BF16) */
  int64_t num_cameras;       /* all cameras (replicated on every rank) */
  int64_t num_points;        /* points owned by this rank */
  int64_t num_observations;  /* observations of those points */
  const int32_t *camera_index; /* host [num_observations] */
  const int32_t *point_index;  /* host [num_observations], local point index */
  int32_t tile_size;         /* max observations per tile, 0 = default (256, the storage tile) */
  int32_t slot_cap;          /* max distinct cameras per super-tile, 0 = default = maximum (192).  Tracks of ANY length are
                              * supported: a point with more observations than min(tile_size, slot_cap) is cut into
                              * fragment tiles whose per-point sums are completed in a second level */
  int64_t super_tile_observations; /* target observations per super-tile (one CTA), 0 = default M / (148*8) */
  int64_t flags;             /* GB_FLAG_* */
} gb_problem_desc;
/* This rank holds a point partition: cameras without a local observation are legal (their sums come from the
 * other ranks through the all-reduce).  Without the flag such a camera is an unused vertex and is rejected. */
#define GB_FLAG_PARTITION 1
/* Build the observation-sized structure tables on the host threads instead of the GPU (the default: per-tile records, slot
 * order, camera-major view by the kernels of csrc/structure_device.cuh; only the greedy cuts stay on the host).  Both give
 * bit-identical arrays; the flag exists for the A/B test and for timing. */
#define GB_FLAG_HOST_TABLES 2

/* Replaces: Graph::initialize_optimization + build_structure (graph.hpp:92-219),
 * FactorDescriptor::initialize_device_ids (factor.hpp:455-467), Hessian::build_structure
 * (hessian.hpp:257-288), SchurComplement::build_structure (schur.hpp:194-225) and
 * PCGSchurSolver::update_structure (solver/pcg_schur.hpp:49-65): sorts the factors by (point, camera),
 * builds the point CSR, the observation tiles and the per-tile camera segments used by the
 * atomic-free reductions. */
int gb_problem_create(gb_context *ctx, const gb_problem_desc *desc, gb_problem **out);
int gb_problem_destroy(gb_problem *p);

/* Host-only view of the same structure build (no GPU needed): used by the CPU test-suite and by callers that
 * want the tiling before committing device memory.  which: 0 cam_idx 1 pt_idx (sorted order) 2 pptr 3 tile_obs
 * 4 tile_pt 5 st_tile 6 st_row 7 row_cam 8 cam_row_ptr 9 cam_row_list 10 slot_of_obs (all int32), 11 rank (uint8),
 * 12 perm (int64; empty when the input was already sorted), 13 ometa (uint32) 14 seg_tab (uint32) 15 pt_tab (uint16)
 * 16 tile meta (8 x int32 per tile: p0, n, np, nseg, seg_off, pt_off, o0, frag), 21 frag_tile 22 hv_pt 23 hv_ptr (int32; long
 * tracks: a point with more observations than one tile holds is cut into fragment tiles - the tiles of the fragments, the
 * points, and each point's range in frag_tile).  out may be NULL to query the count. */
typedef struct gb_structure gb_structure;
int gb_structure_create(const gb_problem_desc *desc, gb_structure **out, char *errbuf, int errlen);
int gb_structure_destroy(gb_structure *s);
int gb_structure_info(const gb_structure *s, int64_t info[12]);
int gb_structure_array(const gb_structure *s, int which, void *out, int64_t *count);
/* the same arrays as they are ON THE DEVICE of a problem (built by the GPU unless GB_FLAG_HOST_TABLES): which = 10
 * slot_of_obs, 13 ometa, 17 packed tile records (bytes), 18 tile_cam, 19 cm_slot, 20 cm_pt (17-20 also answer above) */
int gb_problem_structure_array(gb_problem *p, int which, void *out, int64_t *count);
int gb_structure_hessian(const gb_structure *s, int64_t *colptr, int64_t *rowidx, int64_t *offsets);
int gb_structure_schur(const gb_structure *s, int64_t *colptr, int64_t *rowidx, int64_t *nnz_blocks); /* see gb_schur_structure */

/* Sizes: [0]=n_tiles [1]=n_partial_rows (super-tile x camera) [2]=max_track_length [3]=hessian_dim
 * [4]=n_hessian_blocks [5]=n_hessian_values [6]=device bytes allocated [7]=n_obs [8]=n_super_tiles
 * [9]=n_camera_segments (tile x camera) [10]=storage slots
 * [11]=multi-GPU exchange: 0 single rank, 1 NCCL all-reduce, 2 one-shot all-gather over NVLink peer memory */
int gb_problem_info(const gb_problem *p, int64_t info[12]);

/* Replaces: add_factor(..., obs) (factor.hpp:374-412) / add_vertex (vertex.hpp:240-252) + to_device.
 * Arrays are in the caller's order; element type is T.  cams [n_cams][9] = [w(3) t(3) f k1 k2]
 * (examples/bal.cu:118-125), pts [n_pts][3], obs [n_obs][2]. */
int gb_set_observations(gb_problem *p, const void *obs_host);
int gb_set_vertices(gb_problem *p, const void *cams_host, const void *pts_host);
/* Streaming form of gb_set_observations for callers that feed a new observation batch per solve (sliding-window /
 * incremental use, docs/markdown/main.md "FastLoop"): gb_stage_observations_async starts the host-to-device copy of a
 * PINNED host buffer into staging slot 0 or 1 on a second stream and returns at once, so the copy overlaps the LM
 * iterations running on the previous batch; gb_commit_observations makes the compute stream wait for that slot and
 * installs it (later calls use the new observations).  The host buffer must stay valid until the commit. */
int gb_stage_observations_async(gb_problem *p, const void *obs_pinned_host, int slot);
int gb_commit_observations(gb_problem *p, int slot);
/* Device-pointer forms for callers whose data already lives on the GPU (the reference keeps vertices behind device-visible
 * pointers and observations in managed memory, docs/markdown/memory.md:4-27): same layouts, element type T, copied on the
 * context stream; the call returns after the copy has been enqueued and ordered before later calls. */
int gb_set_observations_device(gb_problem *p, const void *obs_dev);
int gb_set_vertices_device(gb_problem *p, const void *cams_dev /*[n_cams][9]*/, const void *pts_dev /*[n_pts][3]*/);
int gb_get_vertices_device(gb_problem *p, void *cams_dev, void *pts_dev);
/* Replaces the loss_func and precision_matrix arguments of add_factor (factor.hpp:373-412; loss.hpp:15-51):
 *   chi2_f = loss(r^T P r), H += loss' J^T P J, b -= loss' J^T P r (ops/chi2.hpp:9-44, ops/hessian.hpp:58-76).
 * One loss for all factors: GB_LOSS_DEFAULT (identity) or GB_LOSS_HUBER(delta).  precision_host: [n_obs][4] row-major
 * 2x2 per factor in the caller's factor order (element type T), symmetric positive definite; NULL = identity
 * (the reference default, factor.hpp:397-405).  With either set, gb_get_jacobians returns the whitened Jacobians
 * sqrt(loss') U J (P = U^T U) that the library stores. */
typedef enum { GB_LOSS_DEFAULT = 0, GB_LOSS_HUBER = 1 } gb_loss;
int gb_set_loss(gb_problem *p, int loss, double delta);
int gb_set_precision(gb_problem *p, const void *precision_host);
/* Replaces: VertexDescriptor::set_fixed (vertex.hpp:254-266) for cameras and points: cameras_fixed [n_cams] / points_fixed
 * [n_points] (host, != 0 = fixed; NULL = none).  A fixed vertex is never updated and adds no unknowns: the other unknowns get
 * exactly the system the reference builds without its columns (graph.hpp:112-147, ops/linearize.hpp:24-27), factors on fixed
 * vertices still count in chi2.  LAYOUT: fixed vertices KEEP their slot in the vectors of this header (gradient, scales, step,
 * Hessian / Schur exports: zero rows and columns, unit diagonal blocks in S) instead of being renumbered away - use the generic
 * path (graphite_b200_graph.h) for the reference's reduced block order.  Not offered with GB_BF16 storage. */
int gb_set_fixed(gb_problem *p, const uint8_t *cameras_fixed, const uint8_t *points_fixed);
/* User-defined factor: replaces FactorTraits::error / ::jacobian (docs/markdown/main.md:284-289; dispatch in
 * ops/error.hpp:33-96 and ops/linearize.hpp:8-138) for any binary factor of the BAL block shape (vertex 0: 9 parameters,
 * vertex 1: 3 parameters, residual 2) - other camera models, other parameterisations.  The library calls `fn` whenever it
 * needs the factors evaluated at the current vertices; `fn` launches the caller's own kernel on `stream` and returns 0.
 * All pointers are DEVICE memory of element type T.  Jc / Jp are NULL when only residuals are needed (trial-step cost).
 * Loss and precision matrices (gb_set_loss / gb_set_precision) are applied by the library on top, as in the reference
 * (ops/chi2.hpp, ops/hessian.hpp).  fn = NULL restores the built-in BAL reprojection factor. */
typedef struct {
  int64_t num_observations;
  const void *cameras;         /* [n_cams] rows of 10 values: 9 parameters + 1 pad */
  const void *points;          /* [n_points][3] */
  const void *observations;    /* [n_obs][2], caller's factor order */
  const int32_t *camera_index; /* [n_obs], caller's factor order */
  const int32_t *point_index;  /* [n_obs] */
  void *residuals;             /* out [n_obs][2] */
  void *Jc;                    /* out [n_obs][18], column-major 2x9 (ops/linearize.hpp:36-38), or NULL */
  void *Jp;                    /* out [n_obs][6],  column-major 2x3, or NULL */
  void *stream;                /* cudaStream_t the evaluation must be enqueued on */
} gb_factor_eval;
typedef int (*gb_factor_fn)(const gb_factor_eval *eval, void *user);
int gb_set_factor(gb_problem *p, gb_factor_fn fn, void *user);
/* Replaces the in-place update through user pointers (docs/markdown/memory.md:4-13): writes back. */
int gb_get_vertices(gb_problem *p, void *cams_host, void *pts_host);

/* Replaces: Hessian::get_block_col_pointers / row_indices / value_offsets (hessian.hpp:238-248).
 * colptr [n_blocks+1], rowidx/offsets [n_hessian_blocks]; upper-triangular block CSC, (col,row) sorted. */
int gb_hessian_structure(const gb_problem *p, int64_t *colptr, int64_t *rowidx, int64_t *offsets);

/* Replaces: Graph::linearize + Graph::chi2 (graph.hpp:228-290): residuals, analytic Jacobians,
 * Jacobi scales, gradient b = -J~^T r.  chi2 (optional) receives sum r^T r (no 1/2). */
int gb_linearize(gb_problem *p, double *chi2);
/* Replaces: Graph::compute_error + chi2 (graph.hpp:221-234). */
int gb_compute_cost(gb_problem *p, double *chi2);
/* Replaces: Solver::update_values (solver/solver.hpp:20; PCGSchurSolver::update_values, pcg_schur.hpp:67-69) for a caller
 * that linearises ITSELF - the reference's Graph::linearize (graph.hpp:236-290) run by its own levenberg_marquardt with this
 * library plugged in as the Solver.  Imports, from DEVICE memory in the caller's factor order, the residuals [n_obs][2]
 * (FactorDescriptor::residuals) and the Jacobi-SCALED column-major Jacobians J~ = J D, [n_obs][18] and [n_obs][6]
 * (FactorDescriptor::jacobians[0|1].data after scale_jacobians_async, ops/linearize.hpp:140-180).  Requires T == S (as the
 * reference's PCGSchurSolver does).  The imported system is solved as it is: scales = 1, so gb_solve's step is in the
 * caller's scaled space (the Solver::solve contract).  Loss / precision matrices set with gb_set_loss / gb_set_precision
 * are applied on top, as in ops/hessian.hpp:58-76.  A later gb_linearize / gb_lm returns to the built-in linearisation. */
int gb_import_linearization(gb_problem *p, const void *residuals_dev, const void *Jc_dev, const void *Jp_dev);
/* Parity exports (host, element type T): b and scales [hessian_dim] (graph.hpp:57,67); residuals [n_obs][2]
 * in the caller's factor order; Jacobians unscaled, column-major 2x9 / 2x3 per factor (as double). */
int gb_get_gradient(gb_problem *p, void *b_host);
int gb_get_scales(gb_problem *p, void *scales_host);
int gb_get_residuals(gb_problem *p, void *r_host);
int gb_get_jacobians(gb_problem *p, double *Jc_host, double *Jp_host);
/* Replaces: Hessian::update_values + get_values (hessian.hpp:290-307): scaled, undamped J~^T J in the
 * reference's value layout (element type S). */
int gb_hessian_values(gb_problem *p, void *values_host);

/* Linear solver behind Solver<T,S> (solver/solver.hpp:16-24):
 *   GB_SOLVER_PCG_SCHUR  PCGSchurSolver + BlockJacobiSchurPreconditioner (solver/pcg_schur.hpp) — points eliminated
 *   GB_SOLVER_PCG_FULL   PCGSolver + BlockJacobiPreconditioner (solver/pcg.hpp:61-232, preconditioner/block_jacobi.hpp):
 *                        matrix-free PCG on the full camera + point system, the reference's mixed-precision path
 *   GB_SOLVER_DIRECT_SCHUR  EigenSchurLDLTSolver / cudssSchurSolver (solver/eigen_schur.hpp:52-108, solver/cudss_schur.hpp):
 *                        the explicit S factorised - here a dense blocked Cholesky ON THE GPU (S is SPD after damping)
 *                        instead of the reference's host-side SimplicialLDLT - then the same back-substitution; gives the
 *                        exact LM step.  Single rank, 9 n_cams <= 20000.  A factorisation failure makes the solve report
 *                        stop_reason 6 and the LM loop reject the step, as the reference does (eigen_schur.hpp:79-82). */
typedef enum { GB_SOLVER_PCG_SCHUR = 0, GB_SOLVER_PCG_FULL = 1, GB_SOLVER_DIRECT_SCHUR = 2 } gb_solver;

/* Form of the Schur complement the PCG runs on (GB_SOLVER_PCG_SCHUR):
 *   GB_SCHUR_IMPLICIT  matrix-free (B - E W E^T) p per iteration: one streaming pass over the Jacobians
 *   GB_SCHUR_EXPLICIT  S built once per solve (the reference's form, schur.hpp:227-235, ops/schur.hpp:154-188), then a
 *                      block-sparse S p per iteration (schur.hpp:347-393)
 *   GB_SCHUR_AUTO      chosen per problem size and iteration count by the measured rule of DESIGN.md section 3; with
 *                      several ranks always matrix-free (all ranks must run the same form, and S does not shrink with
 *                      the number of ranks)
 * All three give the same iterates up to rounding. */
typedef enum { GB_SCHUR_AUTO = 0, GB_SCHUR_IMPLICIT = 1, GB_SCHUR_EXPLICIT = 2 } gb_schur_mode;

typedef struct {
  int64_t max_iterations;  /* PCGSchurSolver / PCGSolver ctor (pcg_schur.hpp:42-45, pcg.hpp:40-45); bal default 10 */
  double tolerance;        /* 1.0 */
  double rejection_ratio;  /* 5.0 */
  int32_t solver;          /* gb_solver */
  int32_t schur_mode;      /* gb_schur_mode */
} gb_pcg_options;

typedef struct {
  int64_t pcg_iterations;  /* executed */
  double rz_final;
  int32_t stop_reason;     /* 0 max_iter, 1 converged, 2 rejected iterate, 3 rz==0, 4 bad denominator; direct solver: 5 solved,
                            * 6 factorisation failed (solve_ok = false) */
  int32_t schur_mode;      /* the form that ran: GB_SCHUR_IMPLICIT or GB_SCHUR_EXPLICIT */
} gb_solve_info;

/* Replaces: Solver::set_damping_factor (Hessian::apply_damping, hessian.hpp:136-176). */
int gb_set_damping(gb_problem *p, double mu, int use_identity);
/* Replaces: PCGSchurSolver::solve (pcg_schur.hpp:79-168) = SchurComplement::update_values
 * (schur.hpp:227-235, matrix-free here) + BlockJacobiSchurPreconditioner::update_values
 * (block_jacobi_schur.hpp:114-151) + PCG + compute_landmark_update (schur.hpp:279-302).
 * delta_host (optional) receives the scaled-space step (element type T). */
int gb_solve(gb_problem *p, const gb_pcg_options *opt, void *delta_host, gb_solve_info *info);
/* The same with the step written to DEVICE memory (bool Solver::solve(Graph*, T *delta_x, StreamPool&), solver/solver.hpp:24:
 * delta_x is a device vector of hessian_dim values in the scaled space); synchronises the context stream before returning,
 * as PCGSchurSolver::solve does (pcg_schur.hpp:165). */
int gb_solve_device(gb_problem *p, const gb_pcg_options *opt, void *delta_dev, gb_solve_info *info);
/* Parity exports of the reduced system at the current damping (element type T): b_S [9 n_cams]
 * (schur.hpp:237), diagonal blocks of S [n_cams][81] column-major, and y = S x for a host vector. */
int gb_get_schur_rhs(gb_problem *p, void *bS_host);
int gb_get_schur_diagonal(gb_problem *p, void *blocks_host);
int gb_schur_multiply(gb_problem *p, const void *x_host, void *y_host);

/* Replaces: SchurComplement::build_structure (schur.hpp:194-225, 397-585) — the upper-triangular block-CSC of
 * S = B - E C^-1 E^T: block (i, j), i <= j, exists iff i == j or cameras i and j observe a common point; columns and the
 * rows inside a column ascending (csc_utils.hpp:16-50).  colptr [n_cams+1], rowidx [nnz_blocks]; either may be NULL
 * (query nnz_blocks first).  Host-side, cached after the first call. */
int gb_schur_structure(gb_problem *p, int64_t *colptr, int64_t *rowidx, int64_t *nnz_blocks);
/* Replaces: SchurComplement::update_values + get_values with an EXPLICIT S (schur.hpp:227-235, ops/schur.hpp:154-188)
 * at the current damping: values [nnz_blocks][81], column-major 9x9 blocks in the order of gb_schur_structure (element
 * type T).  The same deterministic build the explicit solve mode runs on (one warp per block, tuples in a fixed order). */
int gb_schur_values(gb_problem *p, void *values_host);
/* Replaces: SchurComplement::build_csc_structure + update_csc_values -> csc::build_scalar_csc_structure /
 * update_scalar_csc_values (csc_utils.hpp:73-193): S as a SCALAR upper-triangular CSC matrix of dimension 9 n_cams, the
 * interchange format the reference's direct solvers consume (solver/eigen_schur.hpp:52-108, solver/cudss_schur.hpp): per
 * scalar column the rows of every block of its block column, ascending, cut at the diagonal.  pointers [9 n_cams + 1],
 * indices / values [nnz] (values of element type T); any of them may be NULL (query nnz first). */
int gb_schur_csc(gb_problem *p, int32_t *pointers, int32_t *indices, void *values_host, int64_t *nnz);

/* Replaces: backup_parameters + apply_update + compute_error + chi2 + compute_rho
 * (levenberg_marquardt.hpp:174-185): applies the last solve's step, returns the new chi2 and the rho
 * denominator sum dx(mu dx + b) + 1e-3. */
int gb_try_step(gb_problem *p, double *new_chi2, double *rho_denominator);
/* Replaces: Graph::revert_parameters (graph.hpp:311-318). */
int gb_revert_step(gb_problem *p);

typedef struct {
  double initial_damping;   /* LevenbergMarquardtOptions (levenberg_marquardt.hpp:52-75) */
  int64_t iterations;
  int32_t use_identity;
  int32_t verbose;
  gb_pcg_options pcg;
  const volatile int32_t *stop_flag; /* optional (levenberg_marquardt.hpp:69-70) */
  /* Continuation (bench warm-up / step-wise drivers): resume != 0 keeps the current linearisation instead of
   * re-linearising first; initial_nu > 0 replaces the reference's starting nu = 2. */
  int32_t resume;
  int32_t profile_product;  /* != 0: CUDA-event time every launch of the matrix-free Schur product kernel */
  double initial_nu;
  /* != 0: when the LAST iteration of this call accepts its step, the re-linearisation at the new point
   * (levenberg_marquardt.hpp:192-195) is left to whoever needs it next (the next gb_lm / gb_linearize, or an export) instead
   * of being done before returning.  For callers that run one iteration per call and upload fresh vertices in between,
   * where that linearisation would be computed twice. */
  int32_t defer_final_linearize;
  /* != 0: optimizer::levenberg_marquardt2 (levenberg_marquardt.hpp:255-417), the ORB-SLAM-like early termination: an
   * accepted step that lowers chi2 by less than 0.1 % counts as "bad", three bad accepted steps in a row end the loop. */
  int32_t early_stop;
} gb_lm_options;

typedef struct {
  int64_t iterations;       /* executed */
  double initial_chi2, final_chi2, final_damping;
  int64_t accepted, rejected, pcg_iterations_total;
  double seconds_total;     /* device time of the loop (CUDA events) */
  double seconds_linearize, seconds_prepare, seconds_pcg, seconds_backsubst, seconds_cost;
  double final_nu;
  int64_t product_launches;  /* profile_product: executed launches of the Schur product kernel and their device time */
  double product_seconds;
  double update_seconds;     /* profile_product: device time of the rest of those PCG iterations (reduction, exchange, update) */
  /* why the loop ended (levenberg_marquardt.hpp:224-241): GB_LM_DONE all iterations ran; GB_LM_DAMPING_NOT_FINITE the
   * reference returns false here (gb_lm still returns GB_OK: the vertices hold the last accepted state); GB_LM_RHO_ZERO;
   * GB_LM_STOP_FLAG; GB_LM_EARLY_STOP (levenberg_marquardt2, :403-413) */
  int32_t termination;
  int32_t reserved;
  /* profile_product: where the PCG iterations spent their time, from globaltimer stamps of CTA 0 inside k_pcg_solve,
   * summed over the executed iterations: [0] product phase up to CTA 0's arrival at the grid barrier that ends it,
   * [1] that barrier (waiting for the slowest CTA), [2] scalar sums + multi-GPU exchange (push, flags, waiting for the
   * peers), [3] row sums and vector updates, [4] the second grid barrier, [5] multi-GPU: the part of [2] until this rank's
   * own sums are on their way (the rest of [2] is waiting for the peers) */
  double pcg_phase_seconds[6];
} gb_lm_result;
typedef enum { GB_LM_DONE = 0, GB_LM_DAMPING_NOT_FINITE = 1, GB_LM_RHO_ZERO = 2, GB_LM_STOP_FLAG = 3, GB_LM_EARLY_STOP = 4 } gb_lm_termination;

/* Replaces: optimizer::levenberg_marquardt (levenberg_marquardt.hpp:109-242).  trajectory (optional,
 * host, [iterations][4]) receives initial chi2, current chi2, lambda, executed PCG iterations. */
int gb_lm(gb_problem *p, const gb_lm_options *opt, gb_lm_result *result, double *trajectory);

/* Measurement hooks used by bench.py: number of kernels launched by this context since creation,
 * and event-timed repetitions of one stage (0 linearize, 1 prepare, 2 one PCG iteration, 3 back-subst +
 * update, 4 cost, 5 the Schur product kernel alone, 6 its per-camera reduction alone, 7 the prepare tile kernel alone)
 * -> average milliseconds per repetition. */
int64_t gb_kernel_launches(const gb_context *ctx);
int gb_time_stage(gb_problem *p, int stage, int repetitions, double *ms_avg);

#ifdef __cplusplus
}
#endif
#endif /* GRAPHITE_B200_H */

/* oracle/oracle_bal.h — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * C ABI of the CPU restatement of the reference's Levenberg-Marquardt / Schur / PCG
 * path for bundle adjustment (see oracle_bal.cpp for the file:line map).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library; the product (graphite_b200/) never does.
 *
 * All arrays are caller-owned host memory.  Suffix _f64 / _f32 = graph precision T
 * (S == T, as the reference's Schur path requires: include/graphite/schur.hpp:111-113).
 */
#ifndef ORACLE_BAL_H
#define ORACLE_BAL_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_problem orc_problem;

typedef struct {
  double initial_damping;  /* levenberg_marquardt.hpp:54 (1e-4 in examples/bal.cu:284) */
  int64_t iterations;      /* 50 in examples/bal.cu:289 */
  int64_t pcg_iterations;  /* 10 */
  double pcg_tolerance;    /* 1.0 */
  double rejection_ratio;  /* 5.0 */
  int use_identity;        /* 0 */
  int solver;              /* 0 = PCG on explicit Schur (pcg_schur.hpp), 1 = dense LDL^T on Schur (eigen_schur.hpp),
                              2 = matrix-free PCG on the full system (pcg.hpp) */
  int threads;             /* OpenMP threads; <=0 = all */
} orc_lm_options;

/* cams: [nc][9], pts: [np][3], obs: [m][2] row-major, ids int32; copies everything. */
orc_problem *orc_create_f64(int64_t nc, int64_t np, int64_t m, const int32_t *cam_idx, const int32_t *pt_idx,
                            const double *obs, const double *cams, const double *pts);
orc_problem *orc_create_f32(int64_t nc, int64_t np, int64_t m, const int32_t *cam_idx, const int32_t *pt_idx,
                            const double *obs, const double *cams, const double *pts);
void orc_destroy(orc_problem *);
void orc_set_threads(orc_problem *, int threads);

/* Current parameters (as double, whatever T is). */
void orc_get_params(orc_problem *, double *cams, double *pts);
void orc_set_params(orc_problem *, const double *cams, const double *pts);
/* Loss and per-factor precision matrices (factor.hpp:373-412, loss.hpp:15-51): loss_kind 0 = DefaultLoss,
 * 1 = HuberLoss(delta); P: [m][4] row-major 2x2 per factor, NULL = identity. */
void orc_set_robust(orc_problem *, int loss_kind, double delta, const double *P);

/* ops/error.hpp:250-323 + examples/reprojection_error.cuh:61-99.  r: [m][2] (as double). Returns chi2. */
double orc_residuals(orc_problem *, double *r);
/* Unscaled analytic Jacobians, column-major 2xd per observation (ops/linearize.hpp:36-38). */
void orc_jacobians(orc_problem *, double *Jc /*[m][18]*/, double *Jp /*[m][6]*/);
/* graph.hpp:236-290.  scales,b: [9nc+3np].  Returns chi2. */
double orc_linearize(orc_problem *, double *scales, double *b);
/* hessian.hpp:257-288 + csc_utils.hpp:16-50.  Sizes: colptr nblk+1, rowidx/offsets nc+m+np. */
void orc_hessian_structure(orc_problem *, int64_t *colptr, int64_t *rowidx, int64_t *offsets);
int64_t orc_hessian_num_values(orc_problem *);
/* ops/hessian.hpp:9-78, reference value layout, scaled, undamped; requires orc_linearize. */
void orc_hessian_values(orc_problem *, double *values);
/* hessian.hpp:136-176 then schur.hpp:227-235.  Requires orc_linearize.
 * S: dense [9nc][9nc] column-major, only the block-upper triangle is written (lower blocks zero);
 * bS: [9nc]. */
void orc_schur(orc_problem *, double mu, int use_identity, double *S_dense, double *bS);
int64_t orc_schur_nnz_blocks(orc_problem *);
/* One solve at damping mu (set_damping_factor + solve): delta [9nc+3np] in the scaled space.
 * Returns the number of PCG iterations executed (0 for the direct solver). */
int64_t orc_solve(orc_problem *, const orc_lm_options *, double mu, double *delta);
/* levenberg_marquardt.hpp:109-242.  traj: [iterations][4] = initial chi2, current chi2, lambda, pcg iters.
 * Returns the number of iterations executed. */
int64_t orc_lm(orc_problem *, const orc_lm_options *, double *traj);
/* The same loop one iteration at a time (bench warm-up / timed split): begin = everything before the loop,
 * step = one loop body; out4 = initial chi2, current chi2, lambda, pcg iters; returns 0 when the loop ends. */
void orc_lm_begin(orc_problem *, const orc_lm_options *);
int orc_lm_step(orc_problem *, const orc_lm_options *, double *out4);
/* Stage timings of the last orc_lm/orc_solve in seconds: linearize, hessian, schur, pcg, backsubst, update+cost */
void orc_last_timings(orc_problem *, double *t6);

#ifdef __cplusplus
}
#endif
#endif

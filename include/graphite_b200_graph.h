/* graphite_b200_graph.h — C ABI of the GENERIC factor-graph path: vertex sets of any dimension, factor sets of any arity
 * and residual size (unary priors, pose-graph edges 6/6/6, BAL 9/3/2, n-ary factors), fixed vertices, per-factor activity
 * levels, precision matrices and robust losses, evaluated by the CALLER's kernels and assembled / solved by this library
 * (full-system PCG + block-Jacobi, the reference's PCGSolver path) inside the same LM loop as graphite_b200.h.
 *
 * It is the B200-native counterpart of the reference's templated FactorDescriptor / VertexDescriptor machinery
 * (include/graphite/factor.hpp:120-190,373-412; vertex.hpp:54-384; graph.hpp:92-318); the reference-side binding keeps the
 * templated Traits API and passes device pointers through these entry points ("Generic factors").
 * The BAL-shaped Schur path of graphite_b200.h stays the specialised fast path for camera / point problems.
 *
 * Conventions as in graphite_b200.h: int status, gb_last_error(ctx) for the message, nothing throws, no CPU fallback.
 * T = graph precision, S = stored-Jacobian precision ((F64,F64), (F32,F32), (F64,F32), (F64,BF16)).
 * Jacobian blocks are E x d_i column-major per factor (ops/linearize.hpp:36-38), precision matrices E x E row-major
 * (ops/linearize.hpp:283).  All reductions are gathers over a per-vertex incidence list in a fixed order: no atomics,
 * results are bit-reproducible. */
#ifndef GRAPHITE_B200_GRAPH_H
#define GRAPHITE_B200_GRAPH_H
#include "graphite_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

#define GB_MAX_ARITY 4    /* vertices per factor */
#define GB_MAX_DIM 16     /* tangent dimension of a vertex */
#define GB_MAX_RESIDUAL 8 /* residual size E */

typedef struct gb_graph gb_graph;

/* Replaces: VertexDescriptor<T,S,Traits> (vertex.hpp:54-384): add_vertex(id, ptr, fixed) (:240-252), set_fixed (:254-266),
 * set_eliminate.  dimension = Traits::dimension (Hessian columns per vertex); parameters = values stored per vertex
 * (Traits::parameters; 0 = dimension; > dimension for over-parameterised manifolds, then an update callback is needed). */
typedef struct {
  int32_t dimension;
  int32_t parameters;
  int64_t count;
  const int64_t *global_ids; /* host [count]: unique over the whole graph; fixes the block order (graph.hpp:112-147) */
  const uint8_t *fixed;      /* host [count] or NULL: != 0 = fixed vertex (no Hessian column, never updated) */
  int32_t eliminate;         /* != 0: ordered after the non-eliminated sets (graph.hpp:120-128) */
  int32_t reserved;
} gb_vertex_set_desc;

/* Replaces: FactorDescriptor<T,S,Traits> (factor.hpp:120-190) and add_factor(ids, obs, precision, data, loss)
 * (factor.hpp:373-412).  active[f] is FactorDescriptor::set_active's value: the low 7 bits are the optimisation level at
 * which the factor becomes active, the top bit disables it (active.hpp:11-16). */
typedef struct {
  int32_t residual_dim;             /* E = Traits::dimension */
  int32_t arity;                    /* number of vertices per factor */
  int32_t vertex_set[GB_MAX_ARITY]; /* vertex set of every slot (the VertexDescriptors tuple) */
  int64_t count;
  const int32_t *vertex_index;      /* host [count][arity]: index INSIDE the slot's vertex set */
  const uint8_t *active;            /* host [count] or NULL (all 0: active at every level) */
  int32_t loss;                     /* gb_loss */
  int32_t reserved;
  double loss_delta;
} gb_factor_set_desc;

/* Replaces: FactorTraits::error / ::jacobian (docs/markdown/main.md:284-289; dispatch ops/error.hpp:33-96,
 * ops/linearize.hpp:8-138).  The library calls `fn` whenever it needs the factor set evaluated at the current vertices;
 * `fn` launches the caller's kernel on `stream` and returns 0.  All pointers are DEVICE memory, element type T.
 * Only the factors listed in active_index need to be evaluated (the others are ignored).  A Jacobian slot pointer is
 * NULL when with_jacobians == 0 (trial-step cost).  Observations / per-factor data are the caller's own (`user`). */
typedef struct {
  int32_t factor_set;
  int32_t with_jacobians;
  int64_t num_factors;                 /* all factors of the set */
  int64_t num_active;
  const int32_t *active_index;         /* [num_active] ascending factor indices */
  const int32_t *vertex_index;         /* [num_factors][arity] */
  const void *vertices[GB_MAX_ARITY];  /* per slot: [count][parameters] of that slot's vertex set */
  void *residuals;                     /* out [num_factors][E] */
  void *jacobians[GB_MAX_ARITY];       /* out per slot [num_factors][E * d_slot], column-major E x d */
  void *stream;                        /* cudaStream_t */
} gb_graph_eval;
typedef int (*gb_graph_factor_fn)(const gb_graph_eval *eval, void *user);

/* Replaces: VertexTraits::update(vertex, delta) (docs/markdown/main.md:89-111; ops/update.hpp:9-31) for vertex sets whose
 * update is not `parameters[0..d) += delta`.  delta is [count][dimension] in the UNSCALED space (delta~ * scale), rows of
 * inactive / fixed vertices are zero and `active` (device, 1 = active) marks the vertices to touch. */
typedef struct {
  int32_t vertex_set;
  int32_t reserved;
  int64_t count;
  void *vertices;        /* in/out [count][parameters] */
  const void *delta;     /* [count][dimension] */
  const uint8_t *active; /* [count] */
  void *stream;
} gb_graph_update;
typedef int (*gb_graph_update_fn)(const gb_graph_update *upd, void *user);

int gb_graph_create(gb_context *ctx, int precision_T, int precision_S, gb_graph **out);
int gb_graph_destroy(gb_graph *g);
/* both return the id of the new set (>= 0) or a negative gb_status */
int gb_graph_add_vertex_set(gb_graph *g, const gb_vertex_set_desc *desc);
int gb_graph_add_factor_set(gb_graph *g, const gb_factor_set_desc *desc, gb_graph_factor_fn fn, void *user);
int gb_graph_set_update(gb_graph *g, int vertex_set, gb_graph_update_fn fn, void *user);
/* VertexDescriptor::set_fixed / FactorDescriptor::set_active after creation; take effect at the next gb_graph_initialize */
int gb_graph_set_fixed(gb_graph *g, int vertex_set, const uint8_t *fixed_host);
int gb_graph_set_active(gb_graph *g, int factor_set, const uint8_t *active_host);
/* host [count][parameters] of T */
int gb_graph_set_vertices(gb_graph *g, int vertex_set, const void *values_host);
int gb_graph_get_vertices(gb_graph *g, int vertex_set, void *values_host);
/* device pointer of the set's parameters (the caller's kernels may read it; layout [count][parameters] of T) */
int gb_graph_vertices_device(gb_graph *g, int vertex_set, void **ptr);
/* precision matrices: host [count][E*E] row-major of T (stored in S), NULL = identity (factor.hpp:397-405) */
int gb_graph_set_precision(gb_graph *g, int factor_set, const void *precision_host);
int gb_graph_set_loss(gb_graph *g, int factor_set, int loss, double delta);
/* Graph::scale_system (graph.hpp): 0 disables the Jacobi scaling (scales = 1) */
int gb_graph_set_scaling(gb_graph *g, int enable);

/* Replaces: Graph::initialize_optimization(level) + build_structure (graph.hpp:92-219) and Hessian::build_structure
 * (hessian.hpp:257-288): activity of factors at `level`, unused-vertex deactivation (graph.hpp:171-210), block order,
 * per-vertex incidence lists, upper-triangular block-CSC of J^T J.
 * info: [0] hessian_dim [1] block columns (active vertices) [2] hessian blocks [3] hessian values [4] residual rows
 * (all factors, E * count summed) [5] active factors [6] device bytes [7] reserved */
int gb_graph_initialize(gb_graph *g, int level, int64_t info[8]);
/* hessian column offset of every vertex of a set, -1 = inactive (VertexDescriptor::get_hessian_ids) */
int gb_graph_vertex_columns(gb_graph *g, int vertex_set, int64_t *columns_host);
/* Hessian::get_block_col_pointers / row_indices / value offsets (hessian.hpp:238-248) */
int gb_graph_hessian_structure(gb_graph *g, int64_t *colptr, int64_t *rowidx, int64_t *offsets);

/* Graph::linearize + chi2 (graph.hpp:228-290) */
int gb_graph_linearize(gb_graph *g, double *chi2);
/* Graph::compute_error + chi2 (graph.hpp:221-234) */
int gb_graph_cost(gb_graph *g, double *chi2);
/* Exports (host, element type T unless noted): which =
 *   0 b [hessian_dim]                        1 jacobian scales [hessian_dim]
 *   2 residuals of factor set `set` [count][E]   3 chi2 per factor [count]   4 loss derivative dL per factor [count]
 *   5 + slot: stored (scaled) Jacobians of slot `slot` of factor set `set` [count][E*d] (as T)
 *   16 block diagonal of vertex set `set` [count][d*d] column-major (compute_hessian_block_diagonal_async)
 *   17 scalar diagonal of J~^T dL P J~ [hessian_dim] (compute_hessian_scalar_diagonal_async) */
int gb_graph_get(gb_graph *g, int which, int set, void *out_host);
/* Hessian::update_values + get_values: scaled, undamped, element type S (bf16: exported as T) */
int gb_graph_hessian_values(gb_graph *g, void *values_host);
/* FactorDescriptor::compute_Jv / compute_Jtv over ALL factor sets (ops/product.hpp:51-99, 228-290):
 * y [residual rows] = J~ x, x [hessian_dim];  y [hessian_dim] = J~^T dL P v, v [residual rows] (host, T) */
int gb_graph_jv(gb_graph *g, const void *x_host, void *y_host);
int gb_graph_jtpv(gb_graph *g, const void *v_host, void *y_host);

/* Solver::set_damping_factor + PCGSolver::solve with BlockJacobiPreconditioner (solver/pcg.hpp:61-232,
 * preconditioner/block_jacobi.hpp:79-186); only opt->max_iterations / tolerance / rejection_ratio are read.
 * delta_host (optional): the scaled-space step [hessian_dim]. */
int gb_graph_set_damping(gb_graph *g, double mu, int use_identity);
int gb_graph_solve(gb_graph *g, const gb_pcg_options *opt, void *delta_host, gb_solve_info *info);
/* optimizer::levenberg_marquardt (levenberg_marquardt.hpp:109-242) on the generic graph; same options / result /
 * trajectory layout as gb_lm (the Schur-only fields are ignored). */
int gb_graph_lm(gb_graph *g, const gb_lm_options *opt, gb_lm_result *result, double *trajectory);

/* ---- Solver<T,S> plug-in mode (solver/solver.hpp:12-25): Graphite keeps linearising through the user's traits
 * (Graph::linearize, graph.hpp:236-290); the library solves on Graphite's OWN device buffers, used in place.
 *   gb_graph_bind_linearization  per factor set: jacobians_dev[slot] = FactorDescriptor::jacobians[slot].data (E x d
 *       column-major per factor, element type S, Jacobi-scaled by scale_jacobians_async), loss_derivative_dev =
 *       chi2_derivative [count] (S), precision_dev = precision_matrices [count][E*E] (S; NULL = the set's own).
 *       jacobians_dev = NULL unbinds.  Factor callbacks are never called while a linearisation is bound.
 *   gb_graph_bind_gradient       b = Graph::get_b() [hessian_dim] (T), the right-hand side of the solve
 *   gb_graph_update_values       Solver::update_values: the block-Jacobi blocks from the bound Jacobians
 *   gb_graph_solve_device        bool Solver::solve(Graph*, T *delta_x, StreamPool&): the scaled-space step into DEVICE
 *                                memory; synchronises the context stream before returning */
int gb_graph_bind_linearization(gb_graph *g, int factor_set, const void *const *jacobians_dev, const void *loss_derivative_dev,
                                const void *precision_dev);
int gb_graph_bind_gradient(gb_graph *g, const void *b_dev);
int gb_graph_update_values(gb_graph *g);
int gb_graph_solve_device(gb_graph *g, const gb_pcg_options *opt, void *delta_dev, gb_solve_info *info);

#ifdef __cplusplus
}
#endif
#endif /* GRAPHITE_B200_GRAPH_H */

// graphite_b200_adapter.hpp — the Graphite-side binding of libgraphite_b200.so.
//
// This is the header a Graphite maintainer adds next to include/graphite/solver/pcg_schur.hpp.  It is compiled by nvcc
// inside the USER's translation unit, against the unmodified Graphite headers (sfu-rsl/graphite v0.5.0), and calls the
// GPU only through the C ABI of include/graphite_b200.h.  Two levels:
//
//   graphite::B200SchurSolver<T, S, CameraDescriptor, PointDescriptor, FactorDescriptor>
//       a Solver<T,S> (solver/solver.hpp:12-25) that drops into optimizer::levenberg_marquardt in place of
//       PCGSchurSolver + BlockJacobiSchurPreconditioner: Graphite keeps linearising through the user's FactorTraits
//       (Graph::linearize, graph.hpp:236-290); update_values imports its residuals and Jacobi-scaled Jacobians
//       (gb_import_linearization), solve returns the scaled-space step in the device vector delta_x (gb_solve_device).
//
//   graphite::b200_levenberg_marquardt(...)
//       the whole LM loop on the library's fused path (gb_lm) for the built-in BAL reprojection factor: vertex parameters
//       are gathered through VertexDescriptor::vertices() / VertexTraits::parameters, and written back IN PLACE through
//       VertexTraits::update (docs/markdown/memory.md:4-13).
//
// Scope: one binary factor descriptor of the BAL block shape (vertex 0: 9 parameters, not eliminated; vertex 1: 3
// parameters, eliminated; residual dimension 2), every factor and vertex active, T == S in {double, float}.
// Anything else makes update_structure() report the reason on std::cerr and solve() return false, which
// levenberg_marquardt turns into rejected steps (levenberg_marquardt.hpp:181-183).
//
// oracle/adapter_test.cu (built by oracle/Makefile) compiles this header against /root/reference/include and checks it against the
// reference's own PCGSchurSolver.
#pragma once
#include <cstdint>
#include <iostream>
#include <type_traits>
#include <vector>

#include <graphite/graph.hpp>
#include <graphite/loss.hpp>
#include <graphite/optimizer/levenberg_marquardt.hpp>
#include <graphite/solver/solver.hpp>
#include <graphite/stream.hpp>

#include "graphite_b200.h"

namespace graphite {

namespace b200_detail {
template <typename X> struct dtype_of;
template <> struct dtype_of<double> { static constexpr int value = GB_F64; };
template <> struct dtype_of<float> { static constexpr int value = GB_F32; };

template <typename L> struct loss_of { static constexpr bool supported = false; };
template <typename T, int E> struct loss_of<DefaultLoss<T, E>> {
  static constexpr bool supported = true;
  static int kind(const DefaultLoss<T, E> &) { return GB_LOSS_DEFAULT; }
  static double delta(const DefaultLoss<T, E> &) { return 0.0; }
};
template <typename T, int E> struct loss_of<HuberLoss<T, E>> {
  static constexpr bool supported = true;
  static int kind(const HuberLoss<T, E> &) { return GB_LOSS_HUBER; }
  static double delta(const HuberLoss<T, E> &l) { return (double)l.delta; }
};

// camera / point index of every factor in the library's block order (cameras by Hessian offset, then points), from the
// descriptors' own maps: host_ids (factor.hpp:455-461) and local_to_hessian_offsets (vertex.hpp:75, graph.hpp:131-149)
template <typename T, typename S, typename CamD, typename PtD, typename FacD>
bool factor_indices(Graph<T, S> *graph, CamD *cams, PtD *pts, FacD *factors, std::vector<int32_t> &ci,
                    std::vector<int32_t> &pi, std::string &why) {
  static_assert(FacD::N == 2 && FacD::error_dim == 2, "B200 path: binary factors with a 2-d residual");
  static_assert(CamD::dim == 9 && PtD::dim == 3, "B200 path: vertex 0 has 9 parameters, vertex 1 has 3");
  const size_t M = factors->internal_count(), Nc = cams->count(), Np = pts->count();
  if (factors->active_count() != M) { why = "inactive factors are not supported"; return false; }
  if (cams->get_eliminate() || !pts->get_eliminate()) { why = "cameras must be kept and points eliminated"; return false; }
  if (graph->get_hessian_dimension() != 9 * Nc + 3 * Np) { why = "fixed or unused vertices are not supported"; return false; }
  if (graph->get_factor_descriptors().size() != 1 || graph->get_vertex_descriptors().size() != 2) {
    why = "exactly one factor descriptor over two vertex descriptors";
    return false;
  }
  ci.resize(M);
  pi.resize(M);
  for (size_t f = 0; f < M; f++) {
    const size_t hc = cams->local_to_hessian_offsets[factors->host_ids[2 * f]];
    const size_t hp = pts->local_to_hessian_offsets[factors->host_ids[2 * f + 1]];
    if (hc % 9 || hc >= 9 * Nc || hp < 9 * Nc || (hp - 9 * Nc) % 3) { why = "unexpected Hessian block order"; return false; }
    ci[f] = (int32_t)(hc / 9);
    pi[f] = (int32_t)((hp - 9 * Nc) / 3);
  }
  return true;
}
// per-factor precision matrices of the descriptor (factor.hpp:397-405); the identity default needs nothing
template <typename T, typename S, typename FacD> bool forward_precision(gb_problem *prob, FacD *factors) {
  bool identity = true;
  const size_t M = factors->internal_count();
  for (size_t f = 0; f < M && identity; f++) {
    const S *P = &factors->precision_matrices[4 * f];
    identity = P[0] == S(1) && P[1] == S(0) && P[2] == S(0) && P[3] == S(1);
  }
  if (identity) return true;
  std::vector<T> P(4 * M);
  for (size_t i = 0; i < 4 * M; i++) P[i] = (T)factors->precision_matrices[i];
  return gb_set_precision(prob, P.data()) == GB_OK;
}
} // namespace b200_detail

template <typename T, typename S, typename CamD, typename PtD, typename FacD>
class B200SchurSolver : public Solver<T, S> {
  static_assert(std::is_same<T, S>::value, "B200SchurSolver needs T == S (like PCGSchurSolver's preconditioner path)");

  CamD *cams;
  PtD *pts;
  FacD *factors;
  gb_context *ctx = nullptr;
  gb_problem *prob = nullptr;
  gb_pcg_options opt;
  bool ok = false;
  gb_solve_info last{};

  void report(const char *what) const {
    std::cerr << "B200SchurSolver: " << what << ": " << (ctx ? gb_last_error(ctx) : "no context") << std::endl;
  }

public:
  // same constructor arguments as PCGSchurSolver (pcg_schur.hpp:42-45), the descriptors instead of a preconditioner
  B200SchurSolver(size_t max_iter, T tol, T rejection_ratio, CamD *cameras, PtD *points, FacD *factors, int device = 0)
      : cams(cameras), pts(points), factors(factors) {
    opt.max_iterations = (int64_t)max_iter;
    opt.tolerance = (double)tol;
    opt.rejection_ratio = (double)rejection_ratio;
    opt.solver = GB_SOLVER_PCG_SCHUR;
    opt.schur_mode = GB_SCHUR_AUTO;
    if (gb_context_create(device, &ctx) != GB_OK) ctx = nullptr;
  }
  ~B200SchurSolver() override {
    if (prob) gb_problem_destroy(prob);
    if (ctx) gb_context_destroy(ctx);
  }
  B200SchurSolver(const B200SchurSolver &) = delete;
  B200SchurSolver &operator=(const B200SchurSolver &) = delete;

  const gb_solve_info &last_solve() const { return last; }
  gb_problem *problem() { return prob; }

  // PCGSchurSolver::update_structure (pcg_schur.hpp:49-65): one-time sparsity
  void update_structure(Graph<T, S> *graph, StreamPool &) override {
    ok = false;
    if (!ctx) { std::cerr << "B200SchurSolver: no sm_100 device" << std::endl; return; }
    if (prob) { gb_problem_destroy(prob); prob = nullptr; }
    std::vector<int32_t> ci, pi;
    std::string why;
    if (!b200_detail::factor_indices<T, S>(graph, cams, pts, factors, ci, pi, why)) {
      std::cerr << "B200SchurSolver: " << why << std::endl;
      return;
    }
    gb_problem_desc d{};
    d.precision_T = b200_detail::dtype_of<T>::value;
    d.precision_S = b200_detail::dtype_of<S>::value;
    d.num_cameras = (int64_t)cams->count();
    d.num_points = (int64_t)pts->count();
    d.num_observations = (int64_t)ci.size();
    d.camera_index = ci.data();
    d.point_index = pi.data();
    if (gb_problem_create(ctx, &d, &prob) != GB_OK) { report("gb_problem_create"); prob = nullptr; return; }
    // loss and precision matrices of the descriptor (add_factor arguments, factor.hpp:373-412)
    using LossT = typename FacD::LossType;
    static_assert(b200_detail::loss_of<LossT>::supported, "B200 path: DefaultLoss or HuberLoss");
    if (factors->internal_count() > 0) {
      const LossT &l0 = factors->loss[0];
      if (gb_set_loss(prob, b200_detail::loss_of<LossT>::kind(l0), b200_detail::loss_of<LossT>::delta(l0)) != GB_OK) {
        report("gb_set_loss");
        return;
      }
    }
    if (!b200_detail::forward_precision<T, S>(prob, factors)) { report("gb_set_precision"); return; }
    ok = true;
  }

  // PCGSchurSolver::update_values (pcg_schur.hpp:67-69): take over Graphite's linearisation
  void update_values(Graph<T, S> *, StreamPool &) override {
    if (!ok) return;
    cudaDeviceSynchronize(); // Graphite linearises on its own streams
    if (gb_import_linearization(prob, factors->residuals.data().get(), factors->jacobians[0].data.data().get(),
                                factors->jacobians[1].data.data().get()) != GB_OK) {
      report("gb_import_linearization");
      ok = false;
    }
  }

  // PCGSchurSolver::set_damping_factor (pcg_schur.hpp:71-77)
  void set_damping_factor(Graph<T, S> *, T damping_factor, const bool use_identity, StreamPool &) override {
    if (ok && gb_set_damping(prob, (double)damping_factor, use_identity ? 1 : 0) != GB_OK) report("gb_set_damping");
  }

  // PCGSchurSolver::solve (pcg_schur.hpp:79-168): delta_x is a device vector of the Hessian dimension, scaled space
  bool solve(Graph<T, S> *, T *delta_x, StreamPool &) override {
    if (!ok) return false;
    if (gb_solve_device(prob, &opt, delta_x, &last) != GB_OK) {
      report("gb_solve_device");
      return false;
    }
    return true;
  }
};

// ---- whole loop ------------------------------------------------------------------------------------------------------
namespace b200_detail {
template <typename T, typename VTraits, int D>
__global__ void gather_parameters(typename VTraits::Vertex **v, const size_t *hessian_ids, size_t first, size_t n, T *out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  T p[D];
  VTraits::parameters(*v[i], p);
  const size_t row = (hessian_ids[i] - first) / D; // the library's order = the Hessian block order
  for (int k = 0; k < D; k++) out[row * D + k] = p[k];
}
template <typename T, typename VTraits, int D>
__global__ void scatter_parameters(typename VTraits::Vertex **v, const size_t *hessian_ids, size_t first, size_t n,
                                   const T *before, const T *after) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const size_t row = (hessian_ids[i] - first) / D;
  T d[D];
  for (int k = 0; k < D; k++) d[k] = after[row * D + k] - before[row * D + k];
  VTraits::update(*v[i], d); // in place, through the user's own update rule (ops/update.hpp:9-31)
}
} // namespace b200_detail

// optimizer::levenberg_marquardt (levenberg_marquardt.hpp:109-242) for the built-in BAL reprojection factor
// (examples/bal.cuh:15-89), entirely on the library's fused path.  Returns what the reference returns (false when the
// damping factor stopped being finite), result (optional) carries the library's statistics.
template <typename T, typename S, typename CamD, typename PtD, typename FacD>
bool b200_levenberg_marquardt(Graph<T, S> *graph, CamD *cams, PtD *pts, FacD *factors,
                              optimizer::LevenbergMarquardtOptions<T, S> *options, size_t pcg_iterations, T pcg_tolerance,
                              T rejection_ratio, gb_lm_result *result = nullptr, int device = 0) {
  using Obs = typename FacD::ObservationType;
  static_assert(sizeof(Obs) == 2 * sizeof(T), "the observation must be two values of T (examples/bal.cuh:49)");
  if (!graph->initialize_optimization(options->optimization_level) || !graph->build_structure()) return false;
  std::vector<int32_t> ci, pi;
  std::string why;
  if (!b200_detail::factor_indices<T, S>(graph, cams, pts, factors, ci, pi, why)) {
    std::cerr << "b200_levenberg_marquardt: " << why << std::endl;
    return false;
  }
  gb_context *ctx = nullptr;
  gb_problem *prob = nullptr;
  if (gb_context_create(device, &ctx) != GB_OK) return false;
  auto fail = [&](const char *what) {
    std::cerr << "b200_levenberg_marquardt: " << what << ": " << gb_last_error(ctx) << std::endl;
    if (prob) gb_problem_destroy(prob);
    gb_context_destroy(ctx);
    return false;
  };
  const size_t Nc = cams->count(), Np = pts->count(), M = ci.size();
  gb_problem_desc d{};
  d.precision_T = b200_detail::dtype_of<T>::value;
  d.precision_S = b200_detail::dtype_of<S>::value;
  d.num_cameras = (int64_t)Nc; d.num_points = (int64_t)Np; d.num_observations = (int64_t)M;
  d.camera_index = ci.data(); d.point_index = pi.data();
  if (gb_problem_create(ctx, &d, &prob) != GB_OK) return fail("gb_problem_create");
  using LossT = typename FacD::LossType;
  if (M > 0 && gb_set_loss(prob, b200_detail::loss_of<LossT>::kind(factors->loss[0]),
                           b200_detail::loss_of<LossT>::delta(factors->loss[0])) != GB_OK)
    return fail("gb_set_loss");
  if (!b200_detail::forward_precision<T, S>(prob, factors)) return fail("gb_set_precision");
  thrust::device_vector<T> c0(9 * Nc), c1(9 * Nc), p0(3 * Np), p1(3 * Np);
  const int B = 256;
  b200_detail::gather_parameters<T, typename CamD::Traits, 9><<<(unsigned)((Nc + B - 1) / B), B>>>(
      cams->vertices(), cams->get_hessian_ids(), 0, Nc, c0.data().get());
  b200_detail::gather_parameters<T, typename PtD::Traits, 3><<<(unsigned)((Np + B - 1) / B), B>>>(
      pts->vertices(), pts->get_hessian_ids(), 9 * Nc, Np, p0.data().get());
  cudaDeviceSynchronize();
  // the observations sit in the descriptor's managed vector in factor order (factor.hpp:160)
  if (gb_set_observations_device(prob, factors->device_obs.data().get()) != GB_OK) return fail("gb_set_observations_device");
  if (gb_set_vertices_device(prob, c0.data().get(), p0.data().get()) != GB_OK) return fail("gb_set_vertices_device");
  gb_lm_options o{};
  o.initial_damping = (double)options->initial_damping;
  o.iterations = (int64_t)options->iterations;
  o.use_identity = options->use_identity ? 1 : 0;
  o.verbose = options->verbose ? 1 : 0;
  o.pcg.max_iterations = (int64_t)pcg_iterations;
  o.pcg.tolerance = (double)pcg_tolerance;
  o.pcg.rejection_ratio = (double)rejection_ratio;
  o.pcg.solver = GB_SOLVER_PCG_SCHUR;
  gb_lm_result r{};
  if (gb_lm(prob, &o, &r, nullptr) != GB_OK) return fail("gb_lm");
  if (gb_get_vertices_device(prob, c1.data().get(), p1.data().get()) != GB_OK) return fail("gb_get_vertices_device");
  cudaDeviceSynchronize();
  b200_detail::scatter_parameters<T, typename CamD::Traits, 9><<<(unsigned)((Nc + B - 1) / B), B>>>(
      cams->vertices(), cams->get_hessian_ids(), 0, Nc, c0.data().get(), c1.data().get());
  b200_detail::scatter_parameters<T, typename PtD::Traits, 3><<<(unsigned)((Np + B - 1) / B), B>>>(
      pts->vertices(), pts->get_hessian_ids(), 9 * Nc, Np, p0.data().get(), p1.data().get());
  cudaDeviceSynchronize();
  if (result) *result = r;
  gb_problem_destroy(prob);
  gb_context_destroy(ctx);
  return r.termination != GB_LM_DAMPING_NOT_FINITE;
}

} // namespace graphite

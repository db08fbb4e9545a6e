// graphite_b200_graph_adapter.hpp — the Graphite-side binding of the GENERIC factor-graph path of libgraphite_b200.so.
//
// graphite::B200GraphSolver<T, S> is a Solver<T,S> (solver/solver.hpp:12-25) for ANY Graphite graph: whatever vertex and
// factor descriptors the user has defined with the templated Traits API (unary priors, pose-graph edges, BAL factors,
// n-ary factors; autodiff or manual Jacobians; Huber loss; precision matrices; fixed vertices; activity levels).  It drops
// into optimizer::levenberg_marquardt in place of PCGSolver + BlockJacobiPreconditioner (solver/pcg.hpp,
// preconditioner/block_jacobi.hpp):
//
//     graphite::B200GraphSolver<T, S> solver(max_iter, tol, rejection_ratio);
//     solver.add_factor_descriptor(&between);          // every factor descriptor of the graph, any order
//     solver.add_factor_descriptor(&prior);
//     options.solver = &solver;
//     optimizer::levenberg_marquardt<T, S>(&graph, &options);          // unchanged
//
// Graphite keeps linearising through the user's traits (Graph::linearize, graph.hpp:236-290).  The library is handed
// Graphite's OWN device buffers — stored (Jacobi-scaled) Jacobians, loss derivatives, precision matrices, gradient b — and
// uses them in place (gb_graph_bind_linearization / gb_graph_bind_gradient): no copy, no re-evaluation.  The solve is the
// library's single cooperative block-Jacobi PCG kernel over gather-based (atomic-free) products.
//
// Compiled by nvcc inside the user's translation unit against the unmodified Graphite headers; calls the GPU only through
// the C ABI of include/graphite_b200_graph.h.  oracle/adapter_graph_test.cu builds it against /root/reference/include and runs
// it next to the reference's own PCGSolver on the pose-graph fixture (tests/test_adapter.py).
#pragma once
#include <cstdint>
#include <iostream>
#include <memory>
#include <type_traits>
#include <unordered_map>
#include <vector>

#include <graphite/graph.hpp>
#include <graphite/solver/solver.hpp>
#include <graphite/stream.hpp>

#include "graphite_b200_graph.h"

namespace graphite {

namespace b200_graph_detail {
template <typename X> struct dtype_of;
template <> struct dtype_of<double> { static constexpr int value = GB_F64; };
template <> struct dtype_of<float> { static constexpr int value = GB_F32; };
template <> struct dtype_of<__nv_bfloat16> { static constexpr int value = GB_BF16; };

// the library never evaluates factors in plug-in mode; the descriptor still needs a callback
inline int never_called(const gb_graph_eval *, void *) { return 1; }

template <typename T, typename S> struct FactorBinding {
  virtual ~FactorBinding() {}
  // registers the descriptor's factors as a factor set; vset: vertex descriptor -> vertex set id
  virtual int describe(gb_graph *g, const std::unordered_map<const BaseVertexDescriptor<T, S> *, int> &vset) = 0;
  virtual int bind(gb_graph *g, int fset) = 0;
};

template <typename T, typename S, typename F> struct FactorBindingOf : FactorBinding<T, S> {
  F *f;
  explicit FactorBindingOf(F *f) : f(f) {}
  int describe(gb_graph *g, const std::unordered_map<const BaseVertexDescriptor<T, S> *, int> &vset) override {
    constexpr int N = (int)F::N, E = (int)F::error_dim;
    static_assert(N <= GB_MAX_ARITY && E <= GB_MAX_RESIDUAL, "B200 generic path: arity <= 4, residual size <= 8");
    const size_t M = f->internal_count();
    gb_factor_set_desc d{};
    d.residual_dim = E;
    d.arity = N;
    for (int s = 0; s < N; s++) {
      auto it = vset.find(f->vertex_descriptors[s]);
      if (it == vset.end()) return -1000;
      d.vertex_set[s] = it->second;
    }
    d.count = (int64_t)M;
    // local vertex index of every slot (FactorDescriptor::host_ids, factor.hpp:455-461) and the factors that are active at
    // the level Graphite initialised (active_indices, active.hpp:23-48)
    std::vector<int32_t> idx(M * N);
    for (size_t i = 0; i < M * N; i++) idx[i] = (int32_t)f->host_ids[i];
    std::vector<uint8_t> act(M, 0x80);
    thrust::host_vector<size_t> active = f->active_indices;
    for (size_t i = 0; i < active.size(); i++) act[active[i]] = 0;
    d.vertex_index = idx.data();
    d.active = act.data();
    d.loss = GB_LOSS_DEFAULT; // the loss derivative comes with the bound linearisation (chi2_derivative)
    return gb_graph_add_factor_set(g, &d, never_called, nullptr);
  }
  int bind(gb_graph *g, int fset) override {
    const void *jac[GB_MAX_ARITY] = {nullptr, nullptr, nullptr, nullptr};
    for (size_t s = 0; s < F::N; s++) jac[s] = f->jacobians[s].data.data().get();
    return gb_graph_bind_linearization(g, fset, jac, f->chi2_derivative.data().get(), f->precision_matrices.data().get());
  }
};
} // namespace b200_graph_detail

template <typename T, typename S> class B200GraphSolver : public Solver<T, S> {
  gb_context *ctx = nullptr;
  gb_graph *g = nullptr;
  gb_pcg_options opt{};
  gb_solve_info last{};
  std::vector<std::unique_ptr<b200_graph_detail::FactorBinding<T, S>>> factors;
  std::vector<int> fset_of;
  bool ok = false;

  void report(const char *what) const {
    std::cerr << "B200GraphSolver: " << what << ": " << (ctx ? gb_last_error(ctx) : "no context") << std::endl;
  }

public:
  // same constructor arguments as PCGSolver (solver/pcg.hpp:40-45) without the preconditioner object
  B200GraphSolver(size_t max_iter, T tol, T rejection_ratio, int device = 0) {
    opt.max_iterations = (int64_t)max_iter;
    opt.tolerance = (double)tol;
    opt.rejection_ratio = (double)rejection_ratio;
    opt.solver = GB_SOLVER_PCG_FULL;
    if (gb_context_create(device, &ctx) != GB_OK) ctx = nullptr;
  }
  ~B200GraphSolver() override {
    if (g) gb_graph_destroy(g);
    if (ctx) gb_context_destroy(ctx);
  }
  B200GraphSolver(const B200GraphSolver &) = delete;
  B200GraphSolver &operator=(const B200GraphSolver &) = delete;

  template <typename F> void add_factor_descriptor(F *f) {
    factors.emplace_back(new b200_graph_detail::FactorBindingOf<T, S, F>(f));
  }
  const gb_solve_info &last_solve() const { return last; }
  gb_graph *graph_handle() { return g; }

  // PCGSolver::update_structure (pcg.hpp:47-51): called after Graph::initialize_optimization (levenberg_marquardt.hpp:129-137)
  void update_structure(Graph<T, S> *graph, StreamPool &) override {
    ok = false;
    if (!ctx) { std::cerr << "B200GraphSolver: no sm_100 device" << std::endl; return; }
    if (g) { gb_graph_destroy(g); g = nullptr; }
    if (gb_graph_create(ctx, b200_graph_detail::dtype_of<T>::value, b200_graph_detail::dtype_of<S>::value, &g) != GB_OK) {
      report("gb_graph_create");
      g = nullptr;
      return;
    }
    std::unordered_map<const BaseVertexDescriptor<T, S> *, int> vset;
    for (auto *vd : graph->get_vertex_descriptors()) {
      const size_t n = vd->count();
      std::vector<int64_t> gid(n, 0);
      std::vector<uint8_t> fixed(n, 0);
      for (const auto &entry : vd->get_global_map()) { // global id -> local index (vertex.hpp:342-345)
        gid[entry.second] = (int64_t)entry.first;
        fixed[entry.second] = vd->is_fixed(entry.first) ? 1 : 0;
      }
      gb_vertex_set_desc d{};
      d.dimension = (int32_t)vd->dimension();
      d.parameters = 0;
      d.count = (int64_t)n;
      d.global_ids = gid.data();
      d.fixed = fixed.data();
      d.eliminate = vd->get_eliminate() ? 1 : 0;
      const int id = gb_graph_add_vertex_set(g, &d);
      if (id < 0) { report("gb_graph_add_vertex_set"); return; }
      vset[vd] = id;
    }
    if (factors.size() != graph->get_factor_descriptors().size()) {
      std::cerr << "B200GraphSolver: add_factor_descriptor() every factor descriptor of the graph" << std::endl;
      return;
    }
    fset_of.clear();
    for (auto &fb : factors) {
      const int id = fb->describe(g, vset);
      if (id < 0) { report("gb_graph_add_factor_set"); return; }
      fset_of.push_back(id);
    }
    int64_t info[8];
    if (gb_graph_initialize(g, 0, info) != GB_OK) { report("gb_graph_initialize"); return; }
    if ((size_t)info[0] != graph->get_hessian_dimension()) {
      std::cerr << "B200GraphSolver: Hessian dimension " << info[0] << " != Graphite's " << graph->get_hessian_dimension() << std::endl;
      return;
    }
    ok = true;
  }

  // PCGSolver::update_values (pcg.hpp:53-55): Graphite has linearised; use its buffers in place
  void update_values(Graph<T, S> *graph, StreamPool &) override {
    if (!ok) return;
    cudaDeviceSynchronize(); // Graphite linearises on its own streams
    for (size_t i = 0; i < factors.size(); i++)
      if (factors[i]->bind(g, fset_of[i]) != GB_OK) { report("gb_graph_bind_linearization"); ok = false; return; }
    if (gb_graph_bind_gradient(g, graph->get_b().data().get()) != GB_OK || gb_graph_update_values(g) != GB_OK) {
      report("gb_graph_update_values");
      ok = false;
    }
  }

  void set_damping_factor(Graph<T, S> *, T damping_factor, const bool use_identity, StreamPool &) override {
    if (ok && gb_graph_set_damping(g, (double)damping_factor, use_identity ? 1 : 0) != GB_OK) report("gb_graph_set_damping");
  }

  // PCGSolver::solve (pcg.hpp:61-232): delta_x is a device vector of the Hessian dimension, scaled space
  bool solve(Graph<T, S> *, T *delta_x, StreamPool &) override {
    if (!ok) return false;
    if (gb_graph_solve_device(g, &opt, delta_x, &last) != GB_OK) {
      report("gb_graph_solve_device");
      return false;
    }
    return true;
  }
};

} // namespace graphite

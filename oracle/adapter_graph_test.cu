// oracle/adapter_graph_test.cu — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Compiles include/graphite_b200_graph_adapter.hpp against the UNMODIFIED reference headers (where they lie under
// /root/reference/include) and runs, on the pose-graph fixture (6/6/6 autodiff edges with HuberLoss + precision matrices +
// activity levels, unary priors, fixed vertices) and the protocol of oracle/make_golden_pose.py:
//   REFERENCE    the reference's own levenberg_marquardt with PCGSolver + BlockJacobiPreconditioner
//   B200SOLVER   the reference's own levenberg_marquardt (Graph::linearize through the user's traits, apply_update,
//                compute_rho untouched) with graphite::B200GraphSolver plugged in as the Solver<T,S>
// and, before those, ONE solve of the first linearisation with both solvers: DELTA_REL = |dx_b200 - dx_ref| / |dx_ref|.
// tests/test_adapter.py parses the tables.  Built by oracle/Makefile into oracle/_ref/adapter_graph_test, linked against
// graphite_b200/libgraphite_b200.so (the C ABI is the only way in).
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <iostream>
#include <memory>

#include "ref_pose_traits.cuh"

#include <graphite/optimizer/levenberg_marquardt.hpp>
#include <graphite/preconditioner/block_jacobi.hpp>
#include <graphite/solver/pcg.hpp>
#include <graphite/stream.hpp>

#include "../include/graphite_b200_graph_adapter.hpp"

using namespace graphite;

template <typename FP, typename SP> struct PoseScene {
  Graph<FP, SP> graph;
  managed_vector<Pose6V<FP>> poses;
  Pose6Descriptor<FP, SP> pose_desc;
  Between6<FP, SP> between;
  Prior6<FP, SP> prior;

  explicit PoseScene(const PoseProblem &prob) : poses(prob.n), between(&pose_desc, &pose_desc), prior(&pose_desc) {
    pose_desc.reserve(prob.n);
    graph.add_descriptor(&pose_desc);
    between.reserve(prob.mb);
    graph.add_descriptor(&between);
    prior.reserve(prob.mp);
    graph.add_descriptor(&prior);
    for (int64_t i = 0; i < prob.n; i++) {
      for (int j = 0; j < 6; j++) poses[i].v[j] = (FP)prob.poses[6 * i + j];
      pose_desc.add_vertex((size_t)prob.ids[i], &poses[i]);
    }
    for (int64_t i = 0; i < prob.n; i++)
      if (prob.fixed[i]) pose_desc.set_fixed((size_t)prob.ids[i], true);
    const HuberLoss<FP, 6> huber((FP)prob.huber);
    for (int64_t e = 0; e < prob.mb; e++) {
      Meas6<FP> z;
      for (int j = 0; j < 6; j++) z.v[j] = (FP)prob.bt_meas[6 * e + j];
      SP pm[36];
      for (int j = 0; j < 36; j++) pm[j] = (SP)prob.bt_P[36 * e + j];
      const auto id = between.add_factor({(size_t)prob.ids[prob.bt_idx[2 * e]], (size_t)prob.ids[prob.bt_idx[2 * e + 1]]}, z, pm, Empty{}, huber);
      if (prob.bt_active[e]) between.set_active(id, (uint8_t)prob.bt_active[e]);
    }
    for (int64_t e = 0; e < prob.mp; e++) {
      Meas6<FP> z;
      for (int j = 0; j < 6; j++) z.v[j] = (FP)prob.pr_meas[6 * e + j];
      prior.add_factor({(size_t)prob.ids[prob.pr_idx[e]]}, z);
    }
  }
};

template <typename FP, typename SP>
int run(const PoseProblem &prob, size_t iterations, double lambda, size_t pcg_iter, double pcg_tol, int level) {
  const FP rej = 5.0;
  cudaSetDevice(0);
  StreamPool streams(4);
  // ---- one solve of the first linearisation with both solvers -------------------------------------------------------
  {
    PoseScene<FP, SP> a(prob), b(prob);
    BlockJacobiPreconditioner<FP, SP> pre;
    PCGSolver<FP, SP> ref(pcg_iter, (FP)pcg_tol, rej, &pre);
    B200GraphSolver<FP, SP> mine(pcg_iter, (FP)pcg_tol, rej);
    mine.add_factor_descriptor(&b.between);
    mine.add_factor_descriptor(&b.prior);
    thrust::host_vector<FP> dx[2];
    for (int which = 0; which < 2; which++) {
      Graph<FP, SP> *g = which == 0 ? &a.graph : &b.graph;
      Solver<FP, SP> *s = which == 0 ? static_cast<Solver<FP, SP> *>(&ref) : static_cast<Solver<FP, SP> *>(&mine);
      g->initialize_optimization((uint8_t)level);
      g->build_structure();
      s->update_structure(g, streams);
      g->linearize(streams);
      s->update_values(g, streams);
      thrust::device_vector<FP> d(g->get_hessian_dimension());
      s->set_damping_factor(g, (FP)lambda, false, streams);
      const bool ok = s->solve(g, d.data().get(), streams);
      if (!ok) { printf("SOLVE_FAILED %d\n", which); return 3; }
      dx[which] = d;
    }
    double num = 0.0, den = 0.0;
    for (size_t i = 0; i < dx[0].size(); i++) {
      num += ((double)dx[1][i] - (double)dx[0][i]) * ((double)dx[1][i] - (double)dx[0][i]);
      den += (double)dx[0][i] * (double)dx[0][i];
    }
    printf("DELTA_REL %.6e dim %zu pcg_iterations %ld\n", std::sqrt(num / den), dx[0].size(), (long)mine.last_solve().pcg_iterations);
  }
  auto options_for = [&](Solver<FP, SP> *s) {
    optimizer::LevenbergMarquardtOptions<FP, SP> o;
    o.solver = s;
    o.initial_damping = lambda;
    o.iterations = iterations;
    o.optimization_level = (uint8_t)level;
    o.verbose = true;
    o.streams = &streams;
    o.use_identity = false;
    return o;
  };
  {
    PoseScene<FP, SP> sc(prob);
    BlockJacobiPreconditioner<FP, SP> pre;
    PCGSolver<FP, SP> ref(pcg_iter, (FP)pcg_tol, rej, &pre);
    auto o = options_for(&ref);
    printf("== REFERENCE\n");
    fflush(stdout);
    optimizer::levenberg_marquardt<FP, SP>(&sc.graph, &o);
    std::cout << std::flush;
    printf("FINAL_CHI2 %.17g\n", (double)sc.graph.chi2());
  }
  {
    PoseScene<FP, SP> sc(prob);
    B200GraphSolver<FP, SP> mine(pcg_iter, (FP)pcg_tol, rej);
    mine.add_factor_descriptor(&sc.between);
    mine.add_factor_descriptor(&sc.prior);
    auto o = options_for(&mine);
    printf("== B200SOLVER\n");
    fflush(stdout);
    optimizer::levenberg_marquardt<FP, SP>(&sc.graph, &o);
    std::cout << std::flush;
    printf("FINAL_CHI2 %.17g\n", (double)sc.graph.chi2());
  }
  return 0;
}

int main(int argc, char **argv) {
  std::string file, precision = "FP64-FP64";
  size_t iterations = 12, pcg_iter = 30;
  double lambda = 1e-4, pcg_tol = 1e-10;
  int level = 0;
  for (int i = 1; i < argc; i++) {
    std::string s = argv[i];
    if (s == "--iterations") iterations = atol(argv[++i]);
    else if (s == "--precision") precision = argv[++i];
    else if (s == "--lambda") lambda = atof(argv[++i]);
    else if (s == "--pcg_iterations") pcg_iter = atol(argv[++i]);
    else if (s == "--pcg_tolerance") pcg_tol = atof(argv[++i]);
    else if (s == "--level") level = atoi(argv[++i]);
    else file = s;
  }
  PoseProblem p;
  if (!load_pose_graph(file, p)) { fprintf(stderr, "cannot read %s\n", file.c_str()); return 1; }
  printf("POSE_GRAPH %ld %ld %ld precision=%s level=%d\n", (long)p.n, (long)p.mb, (long)p.mp, precision.c_str(), level);
  if (precision == "FP32-FP32") return run<float, float>(p, iterations, lambda, pcg_iter, pcg_tol, level);
  if (precision == "FP64-FP32") return run<double, float>(p, iterations, lambda, pcg_iter, pcg_tol, level);
  return run<double, double>(p, iterations, lambda, pcg_iter, pcg_tol, level);
}

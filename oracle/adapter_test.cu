// oracle/adapter_test.cu — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Compiles include/graphite_b200_adapter.hpp against the UNMODIFIED reference headers (where they lie under
// /root/reference/include) and runs, on the same problem and protocol (examples/bal.cu:284-309):
//   REFERENCE    the reference's own levenberg_marquardt with PCGSchurSolver + BlockJacobiSchurPreconditioner
//   B200SOLVER   the reference's own levenberg_marquardt (Graph::linearize, apply_update, compute_rho untouched) with
//                graphite::B200SchurSolver plugged in as the Solver<T,S>
//   B200LOOP     graphite::b200_levenberg_marquardt: the library's whole loop over the real Vertex/FactorDescriptor
//                members, results written back in place through VertexTraits::update
// and, before those, ONE solve of the first linearisation with both solvers (the check of tests/schur.cu:340-389,
// PCG-Schur against another solver of the same system): DELTA_REL = |dx_b200 - dx_ref| / |dx_ref|.
// tests/test_adapter.py parses the tables.  Built by oracle/Makefile into oracle/_ref/adapter_test, linked against
// graphite_b200/libgraphite_b200.so (the C ABI is the only way in).
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <iostream>
#include <memory>

#include "ref_bal_traits.cuh"

#include <graphite/optimizer/levenberg_marquardt.hpp>
#include <graphite/preconditioner/block_jacobi_schur.hpp>
#include <graphite/solver/pcg_schur.hpp>
#include <graphite/stream.hpp>

#include "../include/graphite_b200_adapter.hpp"

using namespace graphite;

template <typename FP, typename LossT> struct Scene {
  using SP = FP;
  Graph<FP, SP> graph;
  managed_vector<PtV<FP>> points;
  managed_vector<CamV<FP>> cameras;
  PointDescriptor<FP, SP> point_desc;
  CameraDescriptor<FP, SP> camera_desc;
  ReprojectionError<FP, SP, LossT> r_desc;

  Scene(const Problem &prob, const LossT &loss, bool weights)
      : points(prob.np), cameras(prob.nc), r_desc(&camera_desc, &point_desc) {
    // same registration order as examples/bal.cu:78-90
    point_desc.reserve(prob.np);
    graph.add_vertex_descriptor(&point_desc);
    camera_desc.reserve(prob.nc);
    graph.add_descriptor(&camera_desc);
    r_desc.reserve(prob.m);
    graph.add_descriptor(&r_desc);
    for (int64_t i = 0; i < prob.m; i++) {
      Obs2<FP> o{{(FP)prob.obs[2 * i], (FP)prob.obs[2 * i + 1]}};
      if (weights) {
        const SP pa = (SP)(1.0 + (double)((i * 7) % 11) / 22.0), pb = (SP)(0.8 + (double)((i * 5) % 13) / 26.0);
        const SP pc = (SP)(((double)((i * 3) % 7) - 3.0) / 20.0);
        const SP pm[4] = {pa, pc, pc, pb};
        r_desc.add_factor({(size_t)prob.cam_idx[i], (size_t)prob.pt_idx[i] + (size_t)prob.nc}, o, pm, Empty{}, loss);
      } else {
        r_desc.add_factor({(size_t)prob.cam_idx[i], (size_t)prob.pt_idx[i] + (size_t)prob.nc}, o, nullptr, Empty{}, loss);
      }
    }
    for (int64_t i = 0; i < prob.nc; i++) {
      for (int j = 0; j < 9; j++) cameras[i].v[j] = (FP)prob.cams[9 * i + j];
      camera_desc.add_vertex(i, &cameras[i]);
    }
    for (int64_t i = 0; i < prob.np; i++) {
      for (int j = 0; j < 3; j++) points[i].v[j] = (FP)prob.pts[3 * i + j];
      point_desc.add_vertex(i + prob.nc, &points[i]);
    }
    point_desc.set_eliminate(true);
  }
};

template <typename FP, typename LossT>
int run(const Problem &prob, size_t iterations, double lambda, const LossT &loss, bool weights) {
  using SP = FP;
  using CamD = CameraDescriptor<FP, SP>;
  using PtD = PointDescriptor<FP, SP>;
  using FacD = ReprojectionError<FP, SP, LossT>;
  using B200 = B200SchurSolver<FP, SP, CamD, PtD, FacD>;
  const size_t pcg_iter = 10;
  const FP pcg_tol = 1.0, rej = 5.0;
  cudaSetDevice(0);
  StreamPool streams(8);

  // ---- one solve of the first linearisation with both solvers -------------------------------------------------------
  {
    Scene<FP, LossT> a(prob, loss, weights), b(prob, loss, weights);
    BlockJacobiSchurPreconditioner<FP, SP> pre;
    PCGSchurSolver<FP, SP> ref(pcg_iter, pcg_tol, rej, &pre);
    B200 mine(pcg_iter, pcg_tol, rej, &b.camera_desc, &b.point_desc, &b.r_desc);
    thrust::host_vector<FP> dx[2];
    for (int which = 0; which < 2; which++) {
      Graph<FP, SP> *g = which == 0 ? &a.graph : &b.graph;
      Solver<FP, SP> *s = which == 0 ? static_cast<Solver<FP, SP> *>(&ref) : static_cast<Solver<FP, SP> *>(&mine);
      g->initialize_optimization(0);
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
    o.optimization_level = 0;
    o.verbose = true;
    o.streams = &streams;
    o.use_identity = false;
    return o;
  };
  // ---- the reference's LM with its own solver ------------------------------------------------------------------------
  {
    Scene<FP, LossT> sc(prob, loss, weights);
    BlockJacobiSchurPreconditioner<FP, SP> pre;
    PCGSchurSolver<FP, SP> ref(pcg_iter, pcg_tol, rej, &pre);
    auto o = options_for(&ref);
    printf("== REFERENCE\n");
    fflush(stdout);
    optimizer::levenberg_marquardt<FP, SP>(&sc.graph, &o);
    std::cout << std::flush;
    printf("FINAL_CHI2 %.17g\n", (double)sc.graph.chi2());
  }
  // ---- the reference's LM with the B200 solver plugged in ------------------------------------------------------------
  {
    Scene<FP, LossT> sc(prob, loss, weights);
    B200 mine(pcg_iter, pcg_tol, rej, &sc.camera_desc, &sc.point_desc, &sc.r_desc);
    auto o = options_for(&mine);
    printf("== B200SOLVER\n");
    fflush(stdout);
    auto t0 = std::chrono::steady_clock::now();
    optimizer::levenberg_marquardt<FP, SP>(&sc.graph, &o);
    const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::cout << std::flush;
    printf("FINAL_CHI2 %.17g\nSECONDS %.6f\n", (double)sc.graph.chi2(), el);
  }
  // ---- the library's whole loop over the descriptors; the user's vertices are updated in place ------------------------
  {
    Scene<FP, LossT> sc(prob, loss, weights);
    optimizer::LevenbergMarquardtOptions<FP, SP> o;
    o.initial_damping = lambda;
    o.iterations = iterations;
    o.optimization_level = 0;
    o.verbose = true;
    o.use_identity = false;
    gb_lm_result res{};
    printf("== B200LOOP\n");
    fflush(stdout);
    const bool ok = b200_levenberg_marquardt<FP, SP>(&sc.graph, &sc.camera_desc, &sc.point_desc, &sc.r_desc, &o, pcg_iter, pcg_tol,
                                                     rej, &res);
    fflush(stdout);
    // the cost of the vertices the USER holds, evaluated by the reference's own kernels
    sc.r_desc.compute_error();
    printf("FINAL_CHI2 %.17g\nLIB_FINAL_CHI2 %.17g\nOK %d ACCEPTED %ld REJECTED %ld SECONDS %.6f\n", (double)sc.graph.chi2(),
           res.final_chi2, ok ? 1 : 0, (long)res.accepted, (long)res.rejected, res.seconds_total);
  }
  return 0;
}

int main(int argc, char **argv) {
  std::string file, precision = "FP64-FP64";
  size_t iterations = 20;
  double lambda = 1e-4, huber = 0.0;
  bool weights = false;
  for (int i = 1; i < argc; i++) {
    std::string s = argv[i];
    if (s == "--iterations") iterations = atol(argv[++i]);
    else if (s == "--precision") precision = argv[++i];
    else if (s == "--lambda") lambda = atof(argv[++i]);
    else if (s == "--huber") huber = atof(argv[++i]);
    else if (s == "--weights") weights = true;
    else file = s;
  }
  Problem p;
  if (!load_problem(file, p)) { fprintf(stderr, "cannot read %s\n", file.c_str()); return 1; }
  printf("PROBLEM %ld %ld %ld precision=%s\n", (long)p.nc, (long)p.np, (long)p.m, precision.c_str());
  if (huber > 0.0) return run<double, HuberLoss<double, 2>>(p, iterations, lambda, HuberLoss<double, 2>(huber), weights);
  if (precision == "FP32-FP32") return run<float, DefaultLoss<float, 2>>(p, iterations, lambda, DefaultLoss<float, 2>(), weights);
  return run<double, DefaultLoss<double, 2>>(p, iterations, lambda, DefaultLoss<double, 2>(), weights);
}

// oracle/ref_pose_driver.cu — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Drives the UNMODIFIED reference (sfu-rsl/graphite, headers compiled where they lie under /root/reference/include) on a
// POSE GRAPH: 6-dof vertices, binary between factors 6/6/6 (autodiff, HuberLoss, per-factor precision matrices,
// activity levels), unary priors 6/6 (manual Jacobian), fixed vertices — the generic-factor side of the reference that the
// BAL driver (ref_driver.cu) does not touch.  Its outputs pin oracle/oracle_graph.py and the generic CUDA path
// (graphite_b200/csrc/graph_generic.cu):
//   * per-iteration chi2 / lambda table of the reference's own levenberg_marquardt with PCGSolver + BlockJacobiPreconditioner,
//   * Hessian block-CSC structure and values, b, Jacobi scales, hessian column of every vertex after the first linearisation.
// The only non-reference code here: the vertex / factor traits a user would write (the residual is
// tests/user_factor/pose_residual.cuh, shared with the user kernels of the CUDA tests) and file IO.
//
// Built by oracle/Makefile into oracle/_ref/ref_pose (git-ignored, shipped to the GPU box by gpurun).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <memory>
#include <string>
#include <tuple>
#include <vector>

#include "ref_pose_traits.cuh"

struct Args {
  std::string file, precision = "FP64-FP64", dump;
  double lambda = 1e-4, pcg_tol = 1e-10, rej = 5.0;
  size_t iterations = 12, pcg_iter = 30;
  int level = 0;
};

template <typename V> static void dump_vec(const std::string &path, const V &v) {
  using E = typename V::value_type;
  thrust::host_vector<E> h = v;
  FILE *f = fopen(path.c_str(), "wb");
  fwrite(h.data(), sizeof(E), h.size(), f);
  fclose(f);
}

template <typename FP, typename SP> int run(const Args &a, const PoseProblem &prob) {
  using namespace graphite;
  cudaSetDevice(0);
  Graph<FP, SP> graph;
  managed_vector<Pose6V<FP>> poses(prob.n);

  auto pose_desc = Pose6Descriptor<FP, SP>();
  pose_desc.reserve(prob.n);
  graph.add_descriptor(&pose_desc);
  auto between = Between6<FP, SP>(&pose_desc, &pose_desc);
  between.reserve(prob.mb);
  graph.add_descriptor(&between);
  auto prior = Prior6<FP, SP>(&pose_desc);
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

  StreamPool streams(4);

  if (!a.dump.empty()) {
    graph.initialize_optimization((uint8_t)a.level);
    graph.build_structure();
    Hessian<FP, SP> H;
    H.build_structure(&graph, streams);
    graph.linearize(streams);
    FP chi2 = graph.chi2();
    H.update_values(&graph, streams);
    dump_vec(a.dump + ".H_colptr.u64", H.d_col_pointers);
    dump_vec(a.dump + ".H_rowidx.u64", H.d_row_indices);
    dump_vec(a.dump + ".H_offsets.u64", H.d_offsets);
    if constexpr (!std::is_same<SP, __nv_bfloat16>::value) dump_vec(a.dump + ".H_values.bin", H.d_hessian);
    dump_vec(a.dump + ".b.bin", graph.get_b());
    dump_vec(a.dump + ".scales.bin", graph.get_jacobian_scales());
    // hessian column of every vertex (size_t max for inactive ones)
    {
      std::vector<int64_t> cols(prob.n, -1);
      const auto &gmap = pose_desc.get_global_map();
      std::vector<size_t> hids(prob.n);
      cudaMemcpy(hids.data(), pose_desc.get_hessian_ids(), sizeof(size_t) * prob.n, cudaMemcpyDeviceToHost);
      for (int64_t i = 0; i < prob.n; i++) {
        const size_t local = gmap.at((size_t)prob.ids[i]);
        cols[i] = pose_desc.is_active((size_t)prob.ids[i]) ? (int64_t)hids[local] : -1;
      }
      FILE *f = fopen((a.dump + ".columns.i64").c_str(), "wb");
      fwrite(cols.data(), 8, cols.size(), f);
      fclose(f);
    }
    printf("DUMP chi2 %.17g\n", (double)chi2);
    printf("DUMP dimH %zu\n", graph.get_hessian_dimension());
  }

  BlockJacobiPreconditioner<FP, SP> preconditioner;
  PCGSolver<FP, SP> solver(a.pcg_iter, (FP)a.pcg_tol, (FP)a.rej, &preconditioner);
  optimizer::LevenbergMarquardtOptions<FP, SP> options;
  options.solver = &solver;
  options.initial_damping = a.lambda;
  options.iterations = a.iterations;
  options.optimization_level = (uint8_t)a.level;
  options.verbose = true;
  options.streams = &streams;
  optimizer::levenberg_marquardt<FP, SP>(&graph, &options);
  printf("FINAL_CHI2 %.17g\n", (double)graph.chi2());
  if (!a.dump.empty()) {
    std::vector<double> out(6 * prob.n);
    for (int64_t i = 0; i < prob.n; i++)
      for (int j = 0; j < 6; j++) out[6 * i + j] = (double)poses[i].v[j];
    FILE *f = fopen((a.dump + ".final_poses.f64").c_str(), "wb");
    fwrite(out.data(), 8, out.size(), f);
    fclose(f);
  }
  return 0;
}

int main(int argc, char **argv) {
  Args a;
  for (int i = 1; i < argc; i++) {
    std::string s = argv[i];
    auto next = [&]() { return std::string(argv[++i]); };
    if (s == "--precision") a.precision = next();
    else if (s == "--lambda") a.lambda = atof(next().c_str());
    else if (s == "--iterations") a.iterations = atol(next().c_str());
    else if (s == "--pcg_iterations") a.pcg_iter = atol(next().c_str());
    else if (s == "--pcg_tolerance") a.pcg_tol = atof(next().c_str());
    else if (s == "--rejection_ratio") a.rej = atof(next().c_str());
    else if (s == "--level") a.level = atoi(next().c_str());
    else if (s == "--dump") a.dump = next();
    else a.file = s;
  }
  PoseProblem p;
  if (!load_pose_graph(a.file, p)) { fprintf(stderr, "cannot read %s\n", a.file.c_str()); return 1; }
  printf("POSE_GRAPH %ld %ld %ld precision=%s level=%d\n", (long)p.n, (long)p.mb, (long)p.mp, a.precision.c_str(), a.level);
  if (a.precision == "FP64-FP64") return run<double, double>(a, p);
  if (a.precision == "FP32-FP32") return run<float, float>(a, p);
  if (a.precision == "FP64-FP32") return run<double, float>(a, p);
  fprintf(stderr, "unsupported precision\n");
  return 2;
}

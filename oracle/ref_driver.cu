// oracle/ref_driver.cu — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Drives the UNMODIFIED reference (sfu-rsl/graphite, headers compiled where
// they lie under /root/reference/include) on a synthetic BAL problem so that
// its outputs can pin the CPU oracle and the CUDA product path:
//   * per-iteration chi2 / lambda table printed by the reference's own
//     levenberg_marquardt (optimizer/levenberg_marquardt.hpp:216-221),
//   * Hessian block-CSC structure (hessian.hpp:225-231 public members),
//   * first-linearisation b, Jacobi scales, H values, Schur b_S.
//
// The only non-reference code here is (1) an Eigen-free twin of the BAL
// vertex/factor traits (examples/bal.cuh:15-89) because Eigen is absent from
// this image, and (2) file IO. The analytic Jacobian is the reference's own
// examples/projection_jacobians.cuh, included verbatim through oracle/shim/Eigen.
// The residual follows examples/reprojection_error.cuh:61-99 with
// Eigen::AngleAxis::toRotationMatrix written out.
//
// Built by oracle/Makefile into oracle/_ref/ref_bal (git-ignored, shipped to the
// GPU box by gpurun). Nothing under graphite_b200/ links or executes this.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

#include "ref_bal_traits.cuh"

#include <graphite/optimizer/levenberg_marquardt.hpp>
#include <graphite/preconditioner/block_jacobi.hpp>
#include <graphite/preconditioner/block_jacobi_schur.hpp>
#include <graphite/solver/pcg.hpp>
#include <graphite/solver/pcg_schur.hpp>
#include <graphite/stream.hpp>

struct Args {
  std::string file, solver = "pcg-schur", precision = "FP64-FP64", dump;
  double lambda = 1e-4, pcg_tol = 1.0, rej = 5.0;
  size_t iterations = 50, pcg_iter = 10;
  bool identity = false;
  double huber = 0.0;    // > 0: HuberLoss(delta) on every factor (loss.hpp:27-51)
  bool weights = false;  // per-factor precision matrices, the rational pattern of graphite_b200/synthetic.py:precision_matrices
  int64_t fix_cameras = 0; // VertexDescriptor::set_fixed on the first N cameras (vertex.hpp:254-266)
  int64_t fix_points = 0;  // ... and on every K-th point (0: none)
};

template <typename V> static void dump_vec(const std::string &path, const V &v) {
  using E = typename V::value_type;
  thrust::host_vector<E> h = v;
  FILE *f = fopen(path.c_str(), "wb");
  fwrite(h.data(), sizeof(E), h.size(), f);
  fclose(f);
}

template <typename FP, typename SP, typename LossT = graphite::DefaultLoss<FP, 2>>
int run(const Args &a, const Problem &prob, const LossT &loss = LossT()) {
  using namespace graphite;
  cudaSetDevice(0);
  Graph<FP, SP> graph;
  managed_vector<PtV<FP>> points(prob.np);
  managed_vector<CamV<FP>> cameras(prob.nc);

  // Same registration order as examples/bal.cu:78-90.
  auto point_desc = PointDescriptor<FP, SP>();
  point_desc.reserve(prob.np);
  graph.add_vertex_descriptor(&point_desc);
  auto camera_desc = CameraDescriptor<FP, SP>();
  camera_desc.reserve(prob.nc);
  graph.add_descriptor(&camera_desc);
  auto r_desc = ReprojectionError<FP, SP, LossT>(&camera_desc, &point_desc);
  r_desc.reserve(prob.m);
  graph.add_descriptor(&r_desc);

  for (int64_t i = 0; i < prob.m; i++) {
    Obs2<FP> o{{(FP)prob.obs[2 * i], (FP)prob.obs[2 * i + 1]}};
    if (a.weights) {
      // exactly representable pattern (integer arithmetic + one correctly rounded division per entry), SPD
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
  for (int64_t i = 0; i < a.fix_cameras && i < prob.nc; i++) camera_desc.set_fixed(i, true);
  if (a.fix_points > 0)
    for (int64_t i = 0; i < prob.np; i += a.fix_points) point_desc.set_fixed(i + prob.nc, true);

  StreamPool streams(8);

  if (!a.dump.empty()) {
    if constexpr (std::is_same<FP, SP>::value) {
      // First-linearisation dump, mirroring tests/schur.cu:113-160.
      graph.initialize_optimization(0);
      graph.build_structure();
      Hessian<FP, SP> H;
      SchurComplement<FP, SP> schur(H);
      H.build_structure(&graph, streams);
      schur.build_structure(&graph, streams);
      graph.linearize(streams);
      FP chi2 = graph.chi2();
      H.update_values(&graph, streams);
      dump_vec(a.dump + ".H_colptr.u64", H.d_col_pointers);
      dump_vec(a.dump + ".H_rowidx.u64", H.d_row_indices);
      dump_vec(a.dump + ".H_offsets.u64", H.d_offsets);
      dump_vec(a.dump + ".H_values.bin", H.d_hessian);
      dump_vec(a.dump + ".b.bin", graph.get_b());
      dump_vec(a.dump + ".scales.bin", graph.get_jacobian_scales());
      H.apply_damping(&graph, (FP)a.lambda, a.identity, streams);
      schur.update_values(&graph, streams);
      dump_vec(a.dump + ".bS.bin", schur.get_b_Schur());
      dump_vec(a.dump + ".S_colptr.u64", schur.d_col_pointers);
      dump_vec(a.dump + ".S_rowidx.u64", schur.d_row_indices);
      // Schur scalar upper CSC (values layout in the reference is hash-order).
      CSCMatrix<SP, int32_t> d_S;
      schur.build_csc_structure(&graph, d_S);
      schur.update_csc_values(&graph, d_S);
      dump_vec(a.dump + ".Scsc_ptr.i32", d_S.d_pointers);
      dump_vec(a.dump + ".Scsc_idx.i32", d_S.d_indices);
      dump_vec(a.dump + ".Scsc_val.bin", d_S.d_values);
      printf("DUMP chi2 %.17g\n", (double)chi2);
    }
  }

  BlockJacobiPreconditioner<FP, SP> preconditioner;
  std::unique_ptr<SchurPreconditioner<FP, SP>> schur_preconditioner;
  std::unique_ptr<Solver<FP, SP>> solver_ptr;
  if (a.solver == "pcg") {
    solver_ptr = std::make_unique<PCGSolver<FP, SP>>(a.pcg_iter, a.pcg_tol, a.rej, &preconditioner);
  } else if (a.solver == "pcg-schur") {
    if constexpr (std::is_same<FP, SP>::value) {
      schur_preconditioner = std::make_unique<BlockJacobiSchurPreconditioner<FP, SP>>();
      solver_ptr = std::make_unique<PCGSchurSolver<FP, SP>>(a.pcg_iter, a.pcg_tol, a.rej,
                                                            schur_preconditioner.get());
    }
  }
  if (!solver_ptr) {
    fprintf(stderr, "unsupported solver/precision\n");
    return 2;
  }

  optimizer::LevenbergMarquardtOptions<FP, SP> options;
  options.solver = solver_ptr.get();
  options.initial_damping = a.lambda;
  options.iterations = a.iterations;
  options.optimization_level = 0;
  options.verbose = true;
  options.streams = &streams;
  options.use_identity = a.identity;

  auto t0 = std::chrono::steady_clock::now();
  optimizer::levenberg_marquardt<FP, SP>(&graph, &options);
  double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  printf("TOTAL_SECONDS %.6f\n", el);
  printf("FINAL_CHI2 %.17g\n", (double)graph.chi2());
  if (!a.dump.empty()) {
    std::vector<double> cams(9 * prob.nc), pts(3 * prob.np);
    for (int64_t i = 0; i < prob.nc; i++) for (int j = 0; j < 9; j++) cams[9 * i + j] = cameras[i].v[j];
    for (int64_t i = 0; i < prob.np; i++) for (int j = 0; j < 3; j++) pts[3 * i + j] = points[i].v[j];
    FILE *f = fopen((a.dump + ".final_cams.f64").c_str(), "wb"); fwrite(cams.data(), 8, cams.size(), f); fclose(f);
    f = fopen((a.dump + ".final_pts.f64").c_str(), "wb"); fwrite(pts.data(), 8, pts.size(), f); fclose(f);
  }
  solver_ptr.reset();
  return 0;
}

int main(int argc, char **argv) {
  Args a;
  for (int i = 1; i < argc; i++) {
    std::string s = argv[i];
    auto next = [&]() { return std::string(argv[++i]); };
    if (s == "--solver") a.solver = next();
    else if (s == "--precision") a.precision = next();
    else if (s == "--lambda") a.lambda = atof(next().c_str());
    else if (s == "--iterations") a.iterations = atol(next().c_str());
    else if (s == "--pcg_iterations") a.pcg_iter = atol(next().c_str());
    else if (s == "--pcg_tolerance") a.pcg_tol = atof(next().c_str());
    else if (s == "--rejection_ratio") a.rej = atof(next().c_str());
    else if (s == "--identity_damping") a.identity = true;
    else if (s == "--huber") a.huber = atof(next().c_str());
    else if (s == "--weights") a.weights = true;
    else if (s == "--fix_cameras") a.fix_cameras = atol(next().c_str());
    else if (s == "--fix_points") a.fix_points = atol(next().c_str());
    else if (s == "--dump") a.dump = next();
    else a.file = s;
  }
  Problem p;
  if (!load_problem(a.file, p)) { fprintf(stderr, "cannot read %s\n", a.file.c_str()); return 1; }
  printf("PROBLEM %ld %ld %ld solver=%s precision=%s\n", (long)p.nc, (long)p.np, (long)p.m,
         a.solver.c_str(), a.precision.c_str());
  if (a.huber > 0.0) { // robust runs: FP64 only (one more instantiation of the whole reference stack)
    if (a.precision != "FP64-FP64") { fprintf(stderr, "--huber is built for FP64-FP64 only\n"); return 2; }
    return run<double, double, graphite::HuberLoss<double, 2>>(a, p, graphite::HuberLoss<double, 2>(a.huber));
  }
  if (a.precision == "FP64-FP64") return run<double, double>(a, p);
  if (a.precision == "FP32-FP32") return run<float, float>(a, p);
  if (a.precision == "FP64-FP32") return run<double, float>(a, p);
  if (a.precision == "FP64-BF16") return run<double, __nv_bfloat16>(a, p); // examples/bal.cu:186-236, pcg only (types.hpp:10-19)
  fprintf(stderr, "unsupported precision\n");
  return 2;
}

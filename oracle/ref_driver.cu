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

#include <cuda_bf16.h>
#include <Eigen/Core>
#include <graphite/common.hpp>

#include <projection_jacobians.cuh> // from $(REF)/examples, verbatim

#include <graphite/factor.hpp>
#include <graphite/optimizer/levenberg_marquardt.hpp>
#include <graphite/preconditioner/block_jacobi.hpp>
#include <graphite/preconditioner/block_jacobi_schur.hpp>
#include <graphite/solver/pcg.hpp>
#include <graphite/solver/pcg_schur.hpp>
#include <graphite/stream.hpp>
#include <graphite/vertex.hpp>

namespace graphite {

template <typename T> struct CamV { T v[9]; };
template <typename T> struct PtV { T v[3]; };
template <typename T> struct Obs2 { T v[2]; };

template <typename T> struct PointTraits {
  static constexpr size_t dimension = 3;
  using Vertex = PtV<T>;
  template <typename P>
  d_fn static void parameters(const Vertex &vertex, P *parameters) {
    for (int i = 0; i < 3; i++) parameters[i] = static_cast<P>(vertex.v[i]);
  }
  d_fn static void update(Vertex &vertex, const T *delta) {
    for (int i = 0; i < 3; i++) vertex.v[i] += delta[i];
  }
};

template <typename T> struct CameraTraits {
  static constexpr size_t dimension = 9;
  using State = CamV<T>;
  using Vertex = CamV<T>;
  template <typename P>
  d_fn static void parameters(const Vertex &vertex, P *parameters) {
    for (int i = 0; i < 9; i++) parameters[i] = static_cast<P>(vertex.v[i]);
  }
  d_fn static void update(Vertex &vertex, const T *delta) {
    for (int i = 0; i < 9; i++) vertex.v[i] += delta[i];
  }
  d_fn static State get_state(const Vertex &vertex) { return vertex; }
  d_fn static void set_state(Vertex &vertex, const State &state) { vertex = state; }
};

template <typename T, typename S>
using PointDescriptor = VertexDescriptor<T, S, PointTraits<T>>;
template <typename T, typename S>
using CameraDescriptor = VertexDescriptor<T, S, CameraTraits<T>>;

// examples/reprojection_error.cuh:61-99 with AngleAxis::toRotationMatrix spelled out.
template <typename D, typename T>
__device__ static void bal_residual(const D *cam, const D *X, const Obs2<T> &obs, D *error) {
  D R[9] = {D(1), D(0), D(0), D(0), D(1), D(0), D(0), D(0), D(1)}; // row-major
  const D theta = sqrt(cam[0] * cam[0] + cam[1] * cam[1] + cam[2] * cam[2]);
  if (theta > D(0)) {
    const D ax = cam[0] / theta, ay = cam[1] / theta, az = cam[2] / theta;
    const D s = sin(theta), c = cos(theta);
    const D sx = s * ax, sy = s * ay, sz = s * az;
    const D cx = (D(1) - c) * ax, cy = (D(1) - c) * ay, cz = (D(1) - c) * az;
    D tmp;
    tmp = cx * ay; R[1] = tmp - sz; R[3] = tmp + sz;
    tmp = cx * az; R[2] = tmp + sy; R[6] = tmp - sy;
    tmp = cy * az; R[5] = tmp - sx; R[7] = tmp + sx;
    R[0] = cx * ax + c; R[4] = cy * ay + c; R[8] = cz * az + c;
  }
  const D Px = R[0] * X[0] + R[1] * X[1] + R[2] * X[2] + cam[3];
  const D Py = R[3] * X[0] + R[4] * X[1] + R[5] * X[2] + cam[4];
  const D Pz = R[6] * X[0] + R[7] * X[1] + R[8] * X[2] + cam[5];
  const D px = -Px / Pz, py = -Py / Pz;
  const D r2 = px * px + py * py;
  const D rd = D(1.0) + cam[7] * r2 + cam[8] * r2 * r2;
  error[0] = cam[6] * rd * px - static_cast<D>(obs.v[0]);
  error[1] = cam[6] * rd * py - static_cast<D>(obs.v[1]);
}

template <typename T, typename S, typename LossT = DefaultLoss<T, 2>> struct ReprojectionErrorTraits {
  static constexpr size_t dimension = 2;
  using VertexDescriptors = std::tuple<CameraDescriptor<T, S>, PointDescriptor<T, S>>;
  using Observation = Obs2<T>;
  using Data = Empty;
  using Loss = LossT;
  using Differentiation = DifferentiationMode::Manual;

  template <typename D>
  d_fn static void error(const D *camera, const D *point, const Observation &obs, D *error) {
    bal_residual<D, T>(camera, point, obs, error);
  }

  // examples/reprojection_error.cuh:101-126
  template <typename D, size_t I>
  d_fn static void jacobian(const CamV<T> &camera, const PtV<T> &point,
                            const Observation &obs, D *jacobian) {
    Eigen::Matrix<T, 3, 1> rvec, t, X;
    for (int i = 0; i < 3; i++) {
      rvec(i, 0) = camera.v[i];
      t(i, 0) = camera.v[3 + i];
      X(i, 0) = point.v[i];
    }
    Eigen::Matrix<T, 2, 9> Jc;
    Eigen::Matrix<T, 2, 3> Jp;
    projection_simple<T>(rvec, t, camera.v[6], camera.v[7], camera.v[8], X, Jc, Jp);
    if constexpr (I == 0) {
      for (int i = 0; i < 18; i++) jacobian[i] = static_cast<D>(Jc.d[i]);
    } else {
      for (int i = 0; i < 6; i++) jacobian[i] = static_cast<D>(Jp.d[i]);
    }
  }
};

template <typename T, typename S, typename LossT = DefaultLoss<T, 2>>
using ReprojectionError = FactorDescriptor<T, S, ReprojectionErrorTraits<T, S, LossT>>;

} // namespace graphite

struct Args {
  std::string file, solver = "pcg-schur", precision = "FP64-FP64", dump;
  double lambda = 1e-4, pcg_tol = 1.0, rej = 5.0;
  size_t iterations = 50, pcg_iter = 10;
  bool identity = false;
  double huber = 0.0;    // > 0: HuberLoss(delta) on every factor (loss.hpp:27-51)
  bool weights = false;  // per-factor precision matrices, the rational pattern of graphite_b200/synthetic.py:precision_matrices
};

struct Problem {
  int64_t nc, np, m;
  std::vector<int32_t> cam_idx, pt_idx;
  std::vector<double> obs, cams, pts;
};

// GBAL binary: int64 nc,np,m | int32 cam_idx[m] | int32 pt_idx[m] | f64 obs[2m] | f64 cams[9nc] | f64 pts[3np]
static bool load_problem(const std::string &path, Problem &p) {
  FILE *f = fopen(path.c_str(), "rb");
  if (!f) return false;
  int64_t h[3];
  if (fread(h, 8, 3, f) != 3) return false;
  p.nc = h[0]; p.np = h[1]; p.m = h[2];
  p.cam_idx.resize(p.m); p.pt_idx.resize(p.m); p.obs.resize(2 * p.m);
  p.cams.resize(9 * p.nc); p.pts.resize(3 * p.np);
  bool ok = fread(p.cam_idx.data(), 4, p.m, f) == (size_t)p.m &&
            fread(p.pt_idx.data(), 4, p.m, f) == (size_t)p.m &&
            fread(p.obs.data(), 8, 2 * p.m, f) == (size_t)(2 * p.m) &&
            fread(p.cams.data(), 8, 9 * p.nc, f) == (size_t)(9 * p.nc) &&
            fread(p.pts.data(), 8, 3 * p.np, f) == (size_t)(3 * p.np);
  fclose(f);
  return ok;
}

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

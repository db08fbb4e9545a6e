// oracle/ref_pose_traits.cuh — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// The vertex / factor traits a Graphite USER would write for a 6-dof pose graph (between factors 6/6/6 with autodiff,
// HuberLoss and precision matrices; unary priors 6/6 with a manual Jacobian) and the loader of the GPG1 fixture
// (graphite_b200/synthetic.py: write_pose_graph), shared by oracle/ref_pose_driver.cu (golden generator: the unmodified
// reference) and oracle/adapter_graph_test.cu (the reference with include/graphite_b200_graph_adapter.hpp plugged in).
// The residual itself is tests/user_factor/pose_residual.cuh.
#pragma once
#include <cstdint>
#include <cstdio>
#include <string>
#include <tuple>
#include <vector>

#include <cuda_bf16.h>
#include <Eigen/Core>
#include <graphite/common.hpp>
#include <graphite/factor.hpp>
#include <graphite/vertex.hpp>

#define POSE_FN __host__ __device__ inline
#include "../tests/user_factor/pose_residual.cuh"

#include <graphite/graph.hpp>
#include <graphite/hessian.hpp>
#include <graphite/optimizer/levenberg_marquardt.hpp>
#include <graphite/preconditioner/block_jacobi.hpp>
#include <graphite/solver/pcg.hpp>
#include <graphite/stream.hpp>

namespace graphite {

template <typename T> struct Pose6V { T v[6]; };
template <typename T> struct Meas6 { T v[6]; };

template <typename T> struct Pose6Traits {
  static constexpr size_t dimension = 6;
  using Vertex = Pose6V<T>;
  template <typename P> d_fn static void parameters(const Vertex &vertex, P *parameters) {
    for (int i = 0; i < 6; i++) parameters[i] = static_cast<P>(vertex.v[i]);
  }
  d_fn static void update(Vertex &vertex, const T *delta) {
    for (int i = 0; i < 6; i++) vertex.v[i] += delta[i];
  }
};
template <typename T, typename S> using Pose6Descriptor = VertexDescriptor<T, S, Pose6Traits<T>>;

template <typename T, typename S> struct Between6Traits {
  static constexpr size_t dimension = 6;
  using VertexDescriptors = std::tuple<Pose6Descriptor<T, S>, Pose6Descriptor<T, S>>;
  using Observation = Meas6<T>;
  using Data = Empty;
  using Loss = HuberLoss<T, 6>;
  using Differentiation = DifferentiationMode::Auto;
  template <typename D> d_fn static void error(const D *xi, const D *xj, const Observation &obs, D *error) {
    between6_residual<D, T>(xi, xj, obs.v, error);
  }
};
template <typename T, typename S> using Between6 = FactorDescriptor<T, S, Between6Traits<T, S>>;

template <typename T, typename S> struct Prior6Traits {
  static constexpr size_t dimension = 6;
  using VertexDescriptors = std::tuple<Pose6Descriptor<T, S>>;
  using Observation = Meas6<T>;
  using Data = Empty;
  using Loss = DefaultLoss<T, 6>;
  using Differentiation = DifferentiationMode::Manual;
  template <typename D> d_fn static void error(const D *x, const Observation &obs, D *error) {
    prior6_residual<D, T>(x, obs.v, error);
  }
  template <typename D, size_t I> d_fn static void jacobian(const Pose6V<T> &, const Observation &, D *jacobian) {
    for (int k = 0; k < 36; k++) jacobian[k] = static_cast<D>(k % 7 == 0 ? 1 : 0);
  }
};
template <typename T, typename S> using Prior6 = FactorDescriptor<T, S, Prior6Traits<T, S>>;

} // namespace graphite

struct PoseProblem {
  int64_t n, mb, mp;
  double huber;
  std::vector<int64_t> ids, fixed, bt_idx, bt_active, pr_idx;
  std::vector<double> poses, bt_meas, bt_P, pr_meas;
};
template <typename X> static bool rd(FILE *f, std::vector<X> &v, size_t n) {
  v.resize(n);
  return fread(v.data(), sizeof(X), n, f) == n;
}
static bool load_pose_graph(const std::string &path, PoseProblem &p) {
  FILE *f = fopen(path.c_str(), "rb");
  if (!f) return false;
  int64_t h[3];
  if (fread(h, 8, 3, f) != 3 || fread(&p.huber, 8, 1, f) != 1) return false;
  p.n = h[0]; p.mb = h[1]; p.mp = h[2];
  bool ok = rd(f, p.ids, p.n) && rd(f, p.poses, 6 * p.n) && rd(f, p.fixed, p.n) && rd(f, p.bt_idx, 2 * p.mb) &&
            rd(f, p.bt_meas, 6 * p.mb) && rd(f, p.bt_P, 36 * p.mb) && rd(f, p.bt_active, p.mb) && rd(f, p.pr_idx, p.mp) &&
            rd(f, p.pr_meas, 6 * p.mp);
  fclose(f);
  return ok;
}


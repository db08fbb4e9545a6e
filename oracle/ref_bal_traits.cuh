// oracle/ref_bal_traits.cuh — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Eigen-free twin of the reference's BAL vertex / factor traits (examples/bal.cuh:15-89) and the GBAL problem container,
// shared by oracle/ref_driver.cu (the unmodified reference, golden generator) and oracle/adapter_test.cu (the reference
// with include/graphite_b200_adapter.hpp plugged in).  The analytic Jacobian is the reference's own
// examples/projection_jacobians.cuh, included verbatim through oracle/shim/Eigen; the residual follows
// examples/reprojection_error.cuh:61-99 with Eigen::AngleAxis::toRotationMatrix written out.
#pragma once
#include <cstdint>
#include <cstdio>
#include <string>
#include <tuple>
#include <vector>

#include <cuda_bf16.h>
#include <Eigen/Core>
#include <graphite/common.hpp>

#include <projection_jacobians.cuh> // from $(REF)/examples, verbatim

#include <graphite/factor.hpp>
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

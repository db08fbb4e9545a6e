// bal_math.cuh — per-observation BAL camera model on the device.
//
// Residual: r = f (1 + k1 |p|^2 + k2 |p|^4) p - obs,  p = -(R(w) X + t).xy / (R(w) X + t).z
// (reference: examples/reprojection_error.cuh:61-99; camera = [w(3) t(3) f k1 k2]).
// Jacobians: analytic, column-major 2x9 / 2x3 (reference layout, include/graphite/ops/linearize.hpp:36-38;
// the reference's generated twin is examples/projection_jacobians.cuh).  Derived here by the chain rule:
//   G = dr/dP (2x3),  J_t = G,  J_X = G R,  J_w = G d(RX)/dw,  J_f = d p,  J_k1 = f |p|^2 p,  J_k2 = f |p|^4 p
// with d(RX)/dw from R X = cos(th) X + A (w x X) + B (w.X) w,  A = sin(th)/th,  B = (1-cos th)/th^2.
// theta == 0 gives R = I and zero rotation columns, exactly as the reference's else-branch does
// (projection_jacobians.cuh:200-236).
#pragma once
#include <cuda_runtime.h>

namespace gb {

template <typename T> struct BalObs {
  T r[2];
  T Jc[18]; // column-major 2x9
  T Jp[6];  // column-major 2x3
};

template <typename T> __device__ __forceinline__ void sincos_t(T x, T *s, T *c);
template <> __device__ __forceinline__ void sincos_t<double>(double x, double *s, double *c) { sincos(x, s, c); }
template <> __device__ __forceinline__ void sincos_t<float>(float x, float *s, float *c) { sincosf(x, s, c); }

// Rotation matrix (row-major) of the angle-axis vector w; returns theta.
template <typename T> __device__ __forceinline__ T bal_rotation(const T *w, T *R, T *sn, T *cs) {
  const T t2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  const T theta = sqrt(t2);
  R[0] = R[4] = R[8] = T(1);
  R[1] = R[2] = R[3] = R[5] = R[6] = R[7] = T(0);
  *sn = T(0);
  *cs = T(1);
  if (theta > T(0)) {
    const T it = T(1) / theta;
    const T ax = w[0] * it, ay = w[1] * it, az = w[2] * it;
    T s, c;
    sincos_t<T>(theta, &s, &c);
    *sn = s;
    *cs = c;
    const T sx = s * ax, sy = s * ay, sz = s * az;
    const T cx = (T(1) - c) * ax, cy = (T(1) - c) * ay, cz = (T(1) - c) * az;
    T tmp;
    tmp = cx * ay; R[1] = tmp - sz; R[3] = tmp + sz;
    tmp = cx * az; R[2] = tmp + sy; R[6] = tmp - sy;
    tmp = cy * az; R[5] = tmp - sx; R[7] = tmp + sx;
    R[0] = cx * ax + c; R[4] = cy * ay + c; R[8] = cz * az + c;
  }
  return theta;
}

// ---------------------------------------------------------------------------------------------
// Per-camera precomputation.  Everything in the camera model that does not depend on the point is evaluated
// once per camera and parameter change (one thread per camera) instead of once per observation:
//   cx[0..8] R (row-major), cx[9..11] t, cx[12] f, cx[13] k1, cx[14] k2, cx[15..17] w,
//   cx[18] A = sin(th)/th, cx[19] B = (1-cos th)/th^2, cx[20] A' , cx[21] B', cx[22] theta > 0 ? 1 : 0
// The per-observation functions below then need no sin/cos, square root or division by theta.
// ---------------------------------------------------------------------------------------------
constexpr int CAMX = 24;

template <typename T> __device__ __forceinline__ void bal_cam_precompute(const T *cam, T *cx) {
  T R[9], s, c;
  const T theta = bal_rotation(cam, R, &s, &c);
#pragma unroll
  for (int i = 0; i < 9; i++) cx[i] = R[i];
  cx[9] = cam[3]; cx[10] = cam[4]; cx[11] = cam[5];
  cx[12] = cam[6]; cx[13] = cam[7]; cx[14] = cam[8];
  cx[15] = cam[0]; cx[16] = cam[1]; cx[17] = cam[2];
  T A = T(0), B = T(0), Ap = T(0), Bp = T(0);
  if (theta > T(0)) {
    const T t2 = theta * theta;
    A = s / theta;
    B = (T(1) - c) / t2;
    if (t2 < T(1e-4)) {
      Ap = T(-1.0 / 3.0) + t2 * (T(1.0 / 30.0) - t2 * T(1.0 / 840.0));
      Bp = T(-1.0 / 12.0) + t2 * (T(1.0 / 180.0) - t2 * T(1.0 / 6720.0));
    } else {
      Ap = (c - A) / t2;
      Bp = (A - T(2) * B) / t2;
    }
  }
  cx[18] = A; cx[19] = B; cx[20] = Ap; cx[21] = Bp;
  cx[22] = theta > T(0) ? T(1) : T(0);
  cx[23] = T(0);
}

template <typename T> __device__ __forceinline__ void bal_residual_pre(const T *cx, const T *X, const T *obs, T *r) {
  const T Px = cx[0] * X[0] + cx[1] * X[1] + cx[2] * X[2] + cx[9];
  const T Py = cx[3] * X[0] + cx[4] * X[1] + cx[5] * X[2] + cx[10];
  const T Pz = cx[6] * X[0] + cx[7] * X[1] + cx[8] * X[2] + cx[11];
  const T px = -Px / Pz, py = -Py / Pz;
  const T r2 = px * px + py * py;
  const T rd = T(1) + cx[13] * r2 + cx[14] * r2 * r2;
  r[0] = cx[12] * rd * px - obs[0];
  r[1] = cx[12] * rd * py - obs[1];
}

template <typename T>
__device__ __forceinline__ void bal_residual_jacobian_pre(const T *cx, const T *X, const T *obs, BalObs<T> &out) {
  const T *R = cx;
  const T Px = R[0] * X[0] + R[1] * X[1] + R[2] * X[2] + cx[9];
  const T Py = R[3] * X[0] + R[4] * X[1] + R[5] * X[2] + cx[10];
  const T Pz = R[6] * X[0] + R[7] * X[1] + R[8] * X[2] + cx[11];
  const T iz = T(1) / Pz;
  const T px = -Px * iz, py = -Py * iz;
  const T r2 = px * px + py * py;
  const T f = cx[12], k1 = cx[13], k2 = cx[14];
  const T d = T(1) + k1 * r2 + k2 * r2 * r2;
  {
    const T qx = -Px / Pz, qy = -Py / Pz; // the residual uses the reference's division form
    const T q2 = qx * qx + qy * qy;
    const T rd = T(1) + k1 * q2 + k2 * q2 * q2;
    out.r[0] = f * rd * qx - obs[0];
    out.r[1] = f * rd * qy - obs[1];
  }
  const T e = T(2) * k1 + T(4) * k2 * r2;
  const T mfz = -f * iz;
  T G[6];
  G[0] = mfz * (d + e * px * px);
  G[1] = mfz * (e * px * py);
  G[2] = mfz * px * (d + e * r2);
  G[3] = G[1];
  G[4] = mfz * (d + e * py * py);
  G[5] = mfz * py * (d + e * r2);
  T D[9];
#pragma unroll
  for (int i = 0; i < 9; i++) D[i] = T(0);
  if (cx[22] != T(0)) {
    const T *w = cx + 15;
    const T A = cx[18], B = cx[19], Ap = cx[20], Bp = cx[21];
    const T wx0 = w[1] * X[2] - w[2] * X[1];
    const T wx1 = w[2] * X[0] - w[0] * X[2];
    const T wx2 = w[0] * X[1] - w[1] * X[0];
    const T wX = w[0] * X[0] + w[1] * X[1] + w[2] * X[2];
    const T u0 = -A * X[0] + Ap * wx0 + Bp * wX * w[0];
    const T u1 = -A * X[1] + Ap * wx1 + Bp * wX * w[1];
    const T u2 = -A * X[2] + Ap * wx2 + Bp * wX * w[2];
    D[0] = u0 * w[0] + B * (wX + w[0] * X[0]);
    D[1] = u0 * w[1] + A * X[2] + B * (w[0] * X[1]);
    D[2] = u0 * w[2] - A * X[1] + B * (w[0] * X[2]);
    D[3] = u1 * w[0] - A * X[2] + B * (w[1] * X[0]);
    D[4] = u1 * w[1] + B * (wX + w[1] * X[1]);
    D[5] = u1 * w[2] + A * X[0] + B * (w[1] * X[2]);
    D[6] = u2 * w[0] + A * X[1] + B * (w[2] * X[0]);
    D[7] = u2 * w[1] - A * X[0] + B * (w[2] * X[1]);
    D[8] = u2 * w[2] + B * (wX + w[2] * X[2]);
  }
#pragma unroll
  for (int row = 0; row < 2; row++) {
    const T g0 = G[3 * row], g1 = G[3 * row + 1], g2 = G[3 * row + 2];
#pragma unroll
    for (int k = 0; k < 3; k++) {
      out.Jc[row + 2 * k] = g0 * D[k] + g1 * D[3 + k] + g2 * D[6 + k];
      out.Jp[row + 2 * k] = g0 * R[k] + g1 * R[3 + k] + g2 * R[6 + k];
    }
    out.Jc[row + 6] = g0;
    out.Jc[row + 8] = g1;
    out.Jc[row + 10] = g2;
  }
  out.Jc[12] = d * px;
  out.Jc[13] = d * py;
  out.Jc[14] = f * r2 * px;
  out.Jc[15] = f * r2 * py;
  out.Jc[16] = f * r2 * r2 * px;
  out.Jc[17] = f * r2 * r2 * py;
}

} // namespace gb

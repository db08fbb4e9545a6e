// A USER's factor, written once for every scalar type it is evaluated with: plain numbers, the reference's
// graphite::Dual (autodiff, one direction per pass; oracle/ref_pose_driver.cu) and the multi-direction dual of
// tests/user_factor/graph_factors.cu.  TEST INFRASTRUCTURE: this is what a Graphite user puts into FactorTraits::error.
//
// Pose = [w (angle-axis, 3), t (3)], updated by plain addition like the BAL camera (examples/bal.cuh:25-28).
// Between factor (6/6/6): measurement z = [w_z, t_z];  R_e = R_z^T R_i^T R_j,
//   r[0..3) = log(R_e) = theta / (2 sin theta) * vee(R_e - R_e^T),  theta = acos((tr R_e - 1) / 2)
//   r[3..6) = R_i^T (t_j - t_i) - t_z
// Only + - * /, sqrt, sin, cos, acos and comparisons are used (include/graphite/dual.hpp provides exactly these).
#pragma once

#ifndef POSE_FN
#define POSE_FN __host__ __device__ inline
#endif

template <typename D> POSE_FN void pose_rotation(const D *w, D *R /* row-major 3x3 */) {
  R[0] = D(1); R[1] = D(0); R[2] = D(0);
  R[3] = D(0); R[4] = D(1); R[5] = D(0);
  R[6] = D(0); R[7] = D(0); R[8] = D(1);
  const D theta = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  if (theta > D(0)) {
    const D ax = w[0] / theta, ay = w[1] / theta, az = w[2] / theta;
    const D s = sin(theta), c = cos(theta);
    const D sx = s * ax, sy = s * ay, sz = s * az;
    const D cx = (D(1) - c) * ax, cy = (D(1) - c) * ay, cz = (D(1) - c) * az;
    D tmp;
    tmp = cx * ay; R[1] = tmp - sz; R[3] = tmp + sz;
    tmp = cx * az; R[2] = tmp + sy; R[6] = tmp - sy;
    tmp = cy * az; R[5] = tmp - sx; R[7] = tmp + sx;
    R[0] = cx * ax + c; R[4] = cy * ay + c; R[8] = cz * az + c;
  }
}

template <typename D, typename T> POSE_FN void between6_residual(const D *xi, const D *xj, const T *z, D *r) {
  D Ri[9], Rj[9], Rz[9], M[9], Re[9];
  const D wz[3] = {D(z[0]), D(z[1]), D(z[2])};
  pose_rotation<D>(xi, Ri);
  pose_rotation<D>(xj, Rj);
  pose_rotation<D>(wz, Rz);
  for (int a = 0; a < 3; a++)
    for (int b = 0; b < 3; b++) M[3 * a + b] = Ri[a] * Rj[b] + Ri[3 + a] * Rj[3 + b] + Ri[6 + a] * Rj[6 + b]; // Ri^T Rj
  for (int a = 0; a < 3; a++)
    for (int b = 0; b < 3; b++) Re[3 * a + b] = Rz[a] * M[b] + Rz[3 + a] * M[3 + b] + Rz[6 + a] * M[6 + b]; // Rz^T M
  const D c = (Re[0] + Re[4] + Re[8] - D(1)) / D(2);
  D k = D(0.5);
  if (c < D(1)) {
    const D theta = acos(c);
    k = theta / (D(2) * sin(theta));
  }
  r[0] = k * (Re[7] - Re[5]);
  r[1] = k * (Re[2] - Re[6]);
  r[2] = k * (Re[3] - Re[1]);
  const D d0 = xj[3] - xi[3], d1 = xj[4] - xi[4], d2 = xj[5] - xi[5];
  r[3] = Ri[0] * d0 + Ri[3] * d1 + Ri[6] * d2 - D(z[3]);
  r[4] = Ri[1] * d0 + Ri[4] * d1 + Ri[7] * d2 - D(z[4]);
  r[5] = Ri[2] * d0 + Ri[5] * d1 + Ri[8] * d2 - D(z[5]);
}

// unary prior (6/6): r = x - z
template <typename D, typename T> POSE_FN void prior6_residual(const D *x, const T *z, D *r) {
  for (int a = 0; a < 6; a++) r[a] = x[a] - D(z[a]);
}

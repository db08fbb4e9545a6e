// A USER-DEFINED factor for the gb_set_factor tests: what a Graphite user writes as FactorTraits::error / ::jacobian
// (docs/markdown/main.md:284-289), here the BAL reprojection factor itself evaluated by the user's own kernel through
// the library's public header only (bal_math.cuh is used as the user's camera model).  `user` points to a double scale s:
// the factor is s * (projection - observation), so s = 1 must reproduce the built-in factor bit for bit and s = 2
// quadruples the cost.
#include <cuda_runtime.h>

#include "../../include/graphite_b200.h"
#include "../../graphite_b200/csrc/bal_math.cuh"

template <typename T>
__global__ void k_user_factor(long long n, const T *cams, const T *pts, const T *obs, const int *ci, const int *pi, T scale, T *r,
                              T *Jc, T *Jp) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  T cam[10], cx[gb::CAMX];
  for (int k = 0; k < 10; k++) cam[k] = cams[(long long)ci[i] * 10 + k];
  gb::bal_cam_precompute<T>(cam, cx);
  const T X[3] = {pts[3 * (long long)pi[i]], pts[3 * (long long)pi[i] + 1], pts[3 * (long long)pi[i] + 2]};
  const T ob[2] = {obs[2 * i], obs[2 * i + 1]};
  gb::BalObs<T> B;
  gb::bal_residual_jacobian_pre<T>(cx, X, ob, B);
  r[2 * i] = scale * B.r[0];
  r[2 * i + 1] = scale * B.r[1];
  if (Jc) {
    for (int k = 0; k < 18; k++) Jc[18 * i + k] = scale * B.Jc[k];
    for (int k = 0; k < 6; k++) Jp[6 * i + k] = scale * B.Jp[k];
  }
}

extern "C" int user_bal_factor_f64(const gb_factor_eval *e, void *user) {
  const double scale = user ? *(const double *)user : 1.0;
  const long long n = e->num_observations;
  k_user_factor<double><<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)e->stream>>>(
      n, (const double *)e->cameras, (const double *)e->points, (const double *)e->observations, e->camera_index,
      e->point_index, scale, (double *)e->residuals, (double *)e->Jc, (double *)e->Jp);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}
extern "C" int user_failing_factor(const gb_factor_eval *, void *) { return 7; }

// USER-DEFINED factors for the generic factor-graph tests (include/graphite_b200_graph.h): what a Graphite user writes as
// FactorTraits::error / ::jacobian and VertexTraits::update, compiled against the PUBLIC header only.  TEST INFRASTRUCTURE.
//   * linear factors  r = sum_s A_s v_s - obs  with constant E x d_s blocks: the toy factors of the reference's
//     tests/factor.cu:8-125 (Unary J=[1,0], CoupledUnary J=[2,3], Binary J=([1,2],[3,4]), Residual2 J=I)
//   * SE(2) between factor 3/3/3 with analytic Jacobians and an angle-wrapping update callback
//   * 6-dof between factor 6/6/6 and prior 6/6 (pose_residual.cuh), Jacobians by forward-mode dual numbers
//   * the BAL reprojection factor 2/9/3 (the library's own camera model header used as the user's)
#include <cuda_runtime.h>
#include <type_traits>

#include "../../include/graphite_b200_graph.h"
#include "../../graphite_b200/csrc/bal_math.cuh"

#define POSE_FN __host__ __device__ inline
#include "pose_residual.cuh"

// forward-mode dual number with N directions
template <typename T, int N> struct MDual {
  T v;
  T g[N];
  __host__ __device__ MDual() : v(0) { for (int i = 0; i < N; i++) g[i] = 0; }
  template <typename X, typename = typename std::enable_if<std::is_arithmetic<X>::value>::type>
  __host__ __device__ MDual(X x) : v((T)x) { for (int i = 0; i < N; i++) g[i] = 0; }
};
#define MD_BIN(op, expr_v, expr_g)                                                                  \
  template <typename T, int N> __host__ __device__ inline MDual<T, N> operator op(const MDual<T, N> &a, const MDual<T, N> &b) { \
    MDual<T, N> r;                                                                                  \
    r.v = expr_v;                                                                                   \
    for (int i = 0; i < N; i++) r.g[i] = expr_g;                                                    \
    return r;                                                                                       \
  }
MD_BIN(+, a.v + b.v, a.g[i] + b.g[i])
MD_BIN(-, a.v - b.v, a.g[i] - b.g[i])
MD_BIN(*, a.v * b.v, a.g[i] * b.v + a.v * b.g[i])
MD_BIN(/, a.v / b.v, (a.g[i] * b.v - a.v * b.g[i]) / (b.v * b.v))
template <typename T, int N> __host__ __device__ inline bool operator>(const MDual<T, N> &a, const MDual<T, N> &b) { return a.v > b.v; }
template <typename T, int N> __host__ __device__ inline bool operator<(const MDual<T, N> &a, const MDual<T, N> &b) { return a.v < b.v; }
#define MD_UN(name, expr_v, expr_d)                                                         \
  template <typename T, int N> __host__ __device__ inline MDual<T, N> name(const MDual<T, N> &a) { \
    MDual<T, N> r;                                                                          \
    r.v = expr_v;                                                                           \
    const T d = expr_d;                                                                     \
    for (int i = 0; i < N; i++) r.g[i] = a.g[i] * d;                                        \
    return r;                                                                               \
  }
MD_UN(sqrt, ::sqrt(a.v), T(1) / (T(2) * ::sqrt(a.v)))
MD_UN(sin, ::sin(a.v), ::cos(a.v))
MD_UN(cos, ::cos(a.v), -::sin(a.v))
MD_UN(acos, ::acos(a.v), T(-1) / ::sqrt(T(1) - a.v * a.v))

// ---------------------------------------------------------------------------------------------- linear factors
struct LinearUser {
  int E, arity;
  int d[GB_MAX_ARITY];
  const double *A[GB_MAX_ARITY]; // device, column-major E x d_s
  const double *obs;             // device [num_factors][E]
};
template <typename T>
__global__ void k_linear(gb_graph_eval ev, LinearUser u) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ev.num_active) return;
  const long long f = ev.active_index[i];
  T r[GB_MAX_RESIDUAL];
  for (int a = 0; a < u.E; a++) r[a] = -(T)u.obs[f * u.E + a];
  for (int s = 0; s < u.arity; s++) {
    const T *v = (const T *)ev.vertices[s] + (long long)ev.vertex_index[f * u.arity + s] * u.d[s];
    for (int k = 0; k < u.d[s]; k++)
      for (int a = 0; a < u.E; a++) r[a] += (T)u.A[s][k * u.E + a] * v[k];
    if (ev.with_jacobians) {
      T *J = (T *)ev.jacobians[s] + f * u.E * u.d[s];
      for (int k = 0; k < u.E * u.d[s]; k++) J[k] = (T)u.A[s][k];
    }
  }
  for (int a = 0; a < u.E; a++) ((T *)ev.residuals)[f * u.E + a] = r[a];
}
extern "C" int linear_factor_f64(const gb_graph_eval *ev, void *user) {
  k_linear<double><<<(unsigned)((ev->num_active + 127) / 128), 128, 0, (cudaStream_t)ev->stream>>>(*ev, *(const LinearUser *)user);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}
extern "C" int linear_factor_f32(const gb_graph_eval *ev, void *user) {
  k_linear<float><<<(unsigned)((ev->num_active + 127) / 128), 128, 0, (cudaStream_t)ev->stream>>>(*ev, *(const LinearUser *)user);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

// ---------------------------------------------------------------------------------------------- SE(2) between factor
// pose (x, y, th); z = (dx, dy, dth): r = [R(th_i)^T (t_j - t_i) - (dx, dy), wrap(th_j - th_i - dth)]
__host__ __device__ inline double wrap_pi(double a) {
  const double two_pi = 6.283185307179586476925286766559;
  a -= two_pi * floor((a + 3.14159265358979323846) / two_pi);
  return a;
}
__global__ void k_se2(gb_graph_eval ev, const double *meas) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ev.num_active) return;
  const long long f = ev.active_index[i];
  const double *xi = (const double *)ev.vertices[0] + 3ll * ev.vertex_index[2 * f];
  const double *xj = (const double *)ev.vertices[1] + 3ll * ev.vertex_index[2 * f + 1];
  const double *z = meas + 3 * f;
  const double c = cos(xi[2]), s = sin(xi[2]);
  const double dx = xj[0] - xi[0], dy = xj[1] - xi[1];
  double *r = (double *)ev.residuals + 3 * f;
  r[0] = c * dx + s * dy - z[0];
  r[1] = -s * dx + c * dy - z[1];
  r[2] = wrap_pi(xj[2] - xi[2] - z[2]);
  if (ev.with_jacobians) {
    double *Ji = (double *)ev.jacobians[0] + 9 * f, *Jj = (double *)ev.jacobians[1] + 9 * f; // column-major 3x3
    Ji[0] = -c; Ji[1] = s;  Ji[2] = 0;
    Ji[3] = -s; Ji[4] = -c; Ji[5] = 0;
    Ji[6] = -s * dx + c * dy; Ji[7] = -c * dx - s * dy; Ji[8] = -1;
    Jj[0] = c;  Jj[1] = -s; Jj[2] = 0;
    Jj[3] = s;  Jj[4] = c;  Jj[5] = 0;
    Jj[6] = 0;  Jj[7] = 0;  Jj[8] = 1;
  }
}
extern "C" int se2_between_f64(const gb_graph_eval *ev, void *user) {
  k_se2<<<(unsigned)((ev->num_active + 127) / 128), 128, 0, (cudaStream_t)ev->stream>>>(*ev, (const double *)user);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}
// VertexTraits::update for SE(2): add, then wrap the angle
__global__ void k_se2_update(gb_graph_update u) {
  const long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= u.count || !u.active[v]) return;
  double *x = (double *)u.vertices + 3 * v;
  const double *d = (const double *)u.delta + 3 * v;
  x[0] += d[0];
  x[1] += d[1];
  x[2] = wrap_pi(x[2] + d[2]);
}
extern "C" int se2_update_f64(const gb_graph_update *u, void *) {
  k_se2_update<<<(unsigned)((u->count + 127) / 128), 128, 0, (cudaStream_t)u->stream>>>(*u);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

// ---------------------------------------------------------------------------------------------- 6-dof pose graph
template <typename T>
__global__ void k_between6(gb_graph_eval ev, const double *meas) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ev.num_active) return;
  const long long f = ev.active_index[i];
  const T *xi = (const T *)ev.vertices[0] + 6ll * ev.vertex_index[2 * f];
  const T *xj = (const T *)ev.vertices[1] + 6ll * ev.vertex_index[2 * f + 1];
  T z[6];
  for (int k = 0; k < 6; k++) z[k] = (T)meas[6 * f + k];
  T *r = (T *)ev.residuals + 6 * f;
  if (!ev.with_jacobians) {
    between6_residual<T, T>(xi, xj, z, r);
    return;
  }
  using D = MDual<T, 12>;
  D di[6], dj[6], dr[6];
  for (int k = 0; k < 6; k++) {
    di[k] = D(xi[k]); di[k].g[k] = T(1);
    dj[k] = D(xj[k]); dj[k].g[6 + k] = T(1);
  }
  between6_residual<D, T>(di, dj, z, dr);
  T *Ji = (T *)ev.jacobians[0] + 36 * f, *Jj = (T *)ev.jacobians[1] + 36 * f; // column-major 6x6
  for (int a = 0; a < 6; a++) {
    r[a] = dr[a].v;
    for (int k = 0; k < 6; k++) {
      Ji[k * 6 + a] = dr[a].g[k];
      Jj[k * 6 + a] = dr[a].g[6 + k];
    }
  }
}
extern "C" int between6_f64(const gb_graph_eval *ev, void *user) {
  k_between6<double><<<(unsigned)((ev->num_active + 63) / 64), 64, 0, (cudaStream_t)ev->stream>>>(*ev, (const double *)user);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}
extern "C" int between6_f32(const gb_graph_eval *ev, void *user) {
  k_between6<float><<<(unsigned)((ev->num_active + 63) / 64), 64, 0, (cudaStream_t)ev->stream>>>(*ev, (const double *)user);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}
template <typename T>
__global__ void k_prior6(gb_graph_eval ev, const double *meas) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ev.num_active) return;
  const long long f = ev.active_index[i];
  const T *x = (const T *)ev.vertices[0] + 6ll * ev.vertex_index[f];
  T *r = (T *)ev.residuals + 6 * f;
  for (int a = 0; a < 6; a++) r[a] = x[a] - (T)meas[6 * f + a];
  if (ev.with_jacobians) {
    T *J = (T *)ev.jacobians[0] + 36 * f;
    for (int k = 0; k < 36; k++) J[k] = (k % 7 == 0) ? T(1) : T(0);
  }
}
extern "C" int prior6_f64(const gb_graph_eval *ev, void *user) {
  k_prior6<double><<<(unsigned)((ev->num_active + 127) / 128), 128, 0, (cudaStream_t)ev->stream>>>(*ev, (const double *)user);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}
extern "C" int prior6_f32(const gb_graph_eval *ev, void *user) {
  k_prior6<float><<<(unsigned)((ev->num_active + 127) / 128), 128, 0, (cudaStream_t)ev->stream>>>(*ev, (const double *)user);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

// ---------------------------------------------------------------------------------------------- BAL 2/9/3
// vertex set 0: cameras [w t f k1 k2] (9), vertex set 1: points (3); user = device observations [num_factors][2]
template <typename T>
__global__ void k_bal(gb_graph_eval ev, const double *obs) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ev.num_active) return;
  const long long f = ev.active_index[i];
  const T *c = (const T *)ev.vertices[0] + 9ll * ev.vertex_index[2 * f];
  const T *X = (const T *)ev.vertices[1] + 3ll * ev.vertex_index[2 * f + 1];
  T cam[10], cx[gb::CAMX];
  for (int k = 0; k < 9; k++) cam[k] = c[k];
  cam[9] = 0;
  gb::bal_cam_precompute<T>(cam, cx);
  const T ob[2] = {(T)obs[2 * f], (T)obs[2 * f + 1]};
  gb::BalObs<T> B;
  gb::bal_residual_jacobian_pre<T>(cx, X, ob, B);
  T *r = (T *)ev.residuals + 2 * f;
  r[0] = B.r[0];
  r[1] = B.r[1];
  if (ev.with_jacobians) {
    T *Jc = (T *)ev.jacobians[0] + 18 * f, *Jp = (T *)ev.jacobians[1] + 6 * f;
    for (int k = 0; k < 18; k++) Jc[k] = B.Jc[k];
    for (int k = 0; k < 6; k++) Jp[k] = B.Jp[k];
  }
}
extern "C" int bal_graph_factor_f64(const gb_graph_eval *ev, void *user) {
  k_bal<double><<<(unsigned)((ev->num_active + 127) / 128), 128, 0, (cudaStream_t)ev->stream>>>(*ev, (const double *)user);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}
extern "C" int bal_graph_factor_f32(const gb_graph_eval *ev, void *user) {
  k_bal<float><<<(unsigned)((ev->num_active + 127) / 128), 128, 0, (cudaStream_t)ev->stream>>>(*ev, (const double *)user);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}
extern "C" int failing_graph_factor(const gb_graph_eval *, void *) { return 7; }

// small helpers so that the Python tests need no CUDA binding of their own
extern "C" void *user_device_upload(const void *host, long long bytes) {
  void *d = nullptr;
  if (cudaMalloc(&d, bytes > 0 ? bytes : 1) != cudaSuccess) return nullptr;
  if (bytes > 0 && cudaMemcpy(d, host, bytes, cudaMemcpyHostToDevice) != cudaSuccess) { cudaFree(d); return nullptr; }
  return d;
}
extern "C" void user_device_free(void *d) { cudaFree(d); }

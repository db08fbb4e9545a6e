// context.hpp — the context behind gb_context (stream, error text, launch counter, NCCL binding) and the error macros,
// shared by the translation units of the library (api.cu: BAL path; graph_generic.cu: generic factor graphs).
#pragma once
#include "../../include/graphite_b200.h"

#include <cstdarg>
#include <cstdio>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <string>

// ---- NCCL, bound at run time (the library must load on boxes where only torch's bundled NCCL exists) ----
extern "C" {
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId_gb;
}
namespace gbctx {
struct NcclApi {
  void *handle = nullptr;
  int (*GetUniqueId)(ncclUniqueId_gb *) = nullptr;
  int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId_gb, int) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
  bool load() {
    if (handle) return true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
      handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (handle) break;
    }
    if (!handle) return false;
    GetUniqueId = (int (*)(ncclUniqueId_gb *))dlsym(handle, "ncclGetUniqueId");
    CommInitRank = (int (*)(ncclComm_t *, int, ncclUniqueId_gb, int))dlsym(handle, "ncclCommInitRank");
    AllReduce = (int (*)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(handle, "ncclAllReduce");
    AllGather = (int (*)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t))dlsym(handle, "ncclAllGather");
    CommDestroy = (int (*)(ncclComm_t))dlsym(handle, "ncclCommDestroy");
    GetErrorString = (const char *(*)(int))dlsym(handle, "ncclGetErrorString");
    return GetUniqueId && CommInitRank && AllReduce && CommDestroy;
  }
};
inline NcclApi g_nccl;
constexpr int NCCL_INT8 = 0, NCCL_INT32 = 2, NCCL_FLOAT32 = 7, NCCL_FLOAT64 = 8, NCCL_SUM = 0;
} // namespace gbctx
using namespace gbctx;

struct gb_context {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr; // uploads that overlap the compute stream (gb_stage_observations_async)
  bool owns_stream = true;            // false: the caller's stream (gb_context_create_on_stream)
  std::string err;
  int64_t launches = 0;
  ncclComm_t comm = nullptr;
  int nranks = 1, rank = 0;
  int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    err = buf;
    return code;
  }
};

#define GB_CUDA(ctx, call)                                                                          \
  do {                                                                                              \
    cudaError_t e__ = (call);                                                                       \
    if (e__ != cudaSuccess)                                                                         \
      return (ctx)->fail(GB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)
#define GB_TRY(expr)            \
  do {                          \
    int rc__ = (expr);          \
    if (rc__ != GB_OK) return rc__; \
  } while (0)
#define GB_LAUNCH(ctx) ((ctx)->launches++)


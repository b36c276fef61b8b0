// Shared device/host helpers for libvps_b200.so (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/vps_b200.h"

namespace vps {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define VPS_CHECK_ARG(cond, ...)          \
  do {                                    \
    if (!(cond)) {                        \
      vps::set_error(__VA_ARGS__);        \
      return VPS_E_ARG;                   \
    }                                     \
  } while (0)

#define VPS_CUDA_LAST(name)                                                        \
  do {                                                                             \
    cudaError_t e__ = cudaGetLastError();                                          \
    if (e__ != cudaSuccess) {                                                      \
      vps::set_error("%s: %s", name, cudaGetErrorString(e__));                     \
      return VPS_E_CUDA;                                                           \
    }                                                                              \
    vps::count_launch();                                                           \
  } while (0)

static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---- typed element access (activations are f32 or bf16) ----
template <typename T>
__device__ __forceinline__ float ldf(const T* p);
template <>
__device__ __forceinline__ float ldf<float>(const float* p) { return __ldg(p); }
template <>
__device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat162float(*p);
}
template <typename T>
__device__ __forceinline__ void stf(T* p, float v);
template <>
__device__ __forceinline__ void stf<float>(float* p, float v) { *p = v; }
template <>
__device__ __forceinline__ void stf<__nv_bfloat16>(__nv_bfloat16* p, float v) {
  *p = __float2bfloat16_rn(v);
}

// Device view of vps_tensor.
template <typename T>
struct TV {
  T* p;
  int n, h, w, c, cs;
  __device__ __forceinline__ int64_t off(int b, int y, int x) const {
    return (((int64_t)b * h + y) * w + x) * cs;
  }
};
template <typename T>
static inline TV<T> tv(const vps_tensor& t) {
  TV<T> v;
  v.p = (T*)t.ptr;
  v.n = t.n; v.h = t.h; v.w = t.w; v.c = t.c; v.cs = t.cs;
  return v;
}

// dispatch on (in dtype, out dtype) pairs used by the memory-bound kernels
#define VPS_DISPATCH_T(dtype, T, ...)                         \
  do {                                                        \
    if ((dtype) == VPS_F32) { using T = float; __VA_ARGS__; } \
    else { using T = __nv_bfloat16; __VA_ARGS__; }            \
  } while (0)

__device__ __forceinline__ float apply_act(float v, int act, float slope) {
  if (act == VPS_ACT_RELU) return v > 0.f ? v : 0.f;
  if (act == VPS_ACT_LRELU) return v > 0.f ? v : v * slope;
  if (act == VPS_ACT_SIGMOID) return 1.f / (1.f + expf(-v));
  return v;
}

}  // namespace vps

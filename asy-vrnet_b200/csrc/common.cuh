// Shared device/host helpers for libvrcoc (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/vrcoc.h"

namespace vrcoc {

// ---- error plumbing (thread-local, no global mutable state shared between threads) ----------------------
char* err_buf();
int fail(int code, const char* fmt, ...);
int check_launch(const char* what);

#define VRCOC_REQUIRE(cond, ...)                             \
  do {                                                       \
    if (!(cond)) return ::vrcoc::fail(VRCOC_EINVAL, __VA_ARGS__); \
  } while (0)

// ---- dtype-generic scalar access -------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float ldf(const T* p);
template <> __device__ __forceinline__ float ldf<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat162float(*p);
}
template <typename T> __device__ __forceinline__ void stf(T* p, float v);
template <> __device__ __forceinline__ void stf<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void stf<__nv_bfloat16>(__nv_bfloat16* p, float v) {
  *p = __float2bfloat16_rn(v);
}

// runtime-dtype scalar access (used on cold paths / generic gathers)
__device__ __forceinline__ float ld_any(const void* base, int64_t i, int dtype) {
  return dtype == VRCOC_F32 ? __ldg(reinterpret_cast<const float*>(base) + i)
                            : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(base)[i]);
}
__device__ __forceinline__ void st_any(void* base, int64_t i, int dtype, float v) {
  if (dtype == VRCOC_F32) reinterpret_cast<float*>(base)[i] = v;
  else reinterpret_cast<__nv_bfloat16*>(base)[i] = __float2bfloat16_rn(v);
}

// 8 consecutive elements (16-byte aligned for bf16, 32-byte for fp32)
template <typename T> __device__ __forceinline__ void ld8(const T* p, float (&v)[8]);
template <> __device__ __forceinline__ void ld8<float>(const float* p, float (&v)[8]) {
  float4 a = __ldg(reinterpret_cast<const float4*>(p));
  float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <> __device__ __forceinline__ void ld8<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[8]) {
  uint4 r = __ldg(reinterpret_cast<const uint4*>(p));
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x; v[2 * i + 1] = f.y;
  }
}
template <typename T> __device__ __forceinline__ void st8(T* p, const float (&v)[8]);
template <> __device__ __forceinline__ void st8<float>(float* p, const float (&v)[8]) {
  reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
template <> __device__ __forceinline__ void st8<__nv_bfloat16>(__nv_bfloat16* p, const float (&v)[8]) {
  uint4 r;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = r;
}

__device__ __forceinline__ float sigmoidf_exact(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

__device__ __forceinline__ float apply_act(float y, int act) {
  if (act == VRCOC_ACT_RELU) return fmaxf(y, 0.0f);
  if (act == VRCOC_ACT_GELU) return gelu_erf(y);
  if (act == VRCOC_ACT_SILU) return y * sigmoidf_exact(y);
  if (act == VRCOC_ACT_LRELU) return y > 0.f ? y : 0.1f * y;
  return y;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace vrcoc

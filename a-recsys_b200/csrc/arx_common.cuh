// Shared device/host helpers for libarx_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/arx_b200.h"

#define ARX_FULL_MASK 0xffffffffu

#define ARX_CHECK_LAUNCH()                                   \
  do {                                                       \
    cudaError_t e__ = cudaPeekAtLastError();                 \
    if (e__ != cudaSuccess) return ARX_E_LAUNCH;             \
  } while (0)

static inline int arx_num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

// 128-bit read-only load that does not pollute L1 (rows are touched once per CTA).
__device__ __forceinline__ float4 ldg_f4(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}
__device__ __forceinline__ float4 ld_f4(const float* p) {
  return *reinterpret_cast<const float4*>(p);
}
__device__ __forceinline__ void st_f4(float* p, float4 v) {
  *reinterpret_cast<float4*>(p) = v;
}
__device__ __forceinline__ float4 f4_zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ void f4_add(float4& a, const float4 b) {
  a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
}
__device__ __forceinline__ void f4_fma(float4& a, const float w, const float4 b) {
  a.x = fmaf(w, b.x, a.x); a.y = fmaf(w, b.y, a.y);
  a.z = fmaf(w, b.z, a.z); a.w = fmaf(w, b.w, a.w);
}
__device__ __forceinline__ float4 f4_scale(const float4 a, const float s) {
  return make_float4(a.x * s, a.y * s, a.z * s, a.w * s);
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(ARX_FULL_MASK, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(ARX_FULL_MASK, v, o));
  return v;
}

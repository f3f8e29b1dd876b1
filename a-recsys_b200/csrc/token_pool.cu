// Non-linear attribute pooling of the catalog scores (K3m): attributes/embed_attribute.py:194-200.
//   output_feat 2 : logits_f[v, b] = max_{t in bag_f(v)} s_f[t, b]                          (tf.segment_max)
//   output_feat 3 : logits_f[v, b] = m + log(1 + sum_{t in bag_f(v)} exp(s_f[t, b] - m)),  m = max of the WHOLE s_f matrix
// with s_f[t, b] = E_f[t] . u_b + beta_f[t] the score of attribute token t for batch row b.  These poolings are not linear
// in the table rows, so the "pool the catalog first, then contract" rewrite of the default path does not apply: the token
// scores are materialised ([V_f, rows], token-major as in the reference, one tensor-core contraction per attribute) and
// pooled per catalog item here; a categorical attribute is a bag of one token (:172).  The adjoint scatters back to the
// token scores (argmax token / softmax weights, plus the gradient that reaches the global maximum m through the "1 +").
#include "arx_common.cuh"
#include <math_constants.h>

namespace {

enum { TP_GATHER = 1, TP_MAX = 2, TP_LSE = 3 };

__device__ __forceinline__ unsigned long long pack_max(float v, long long idx) {
  unsigned int u = __float_as_uint(v);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);                 // monotone key
  return ((unsigned long long)u << 32) | (unsigned long long)(0xffffffffu - (unsigned int)idx);   // ties -> lowest index
}

// packed (max value, index) of s[t, b] = S[t, b] + bias[t] over the whole [Vf, mb] matrix (Vf * mb < 2^32)
__global__ void score_max_kernel(const float* __restrict__ S, const float* __restrict__ bias, long long Vf, long long mb,
                                 unsigned long long* __restrict__ packed) {
  unsigned long long best = 0ull;
  const long long n = Vf * mb;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = S[i] + (bias ? __ldg(bias + i / mb) : 0.f);
    const unsigned long long p = pack_max(v, i);
    best = p > best ? p : best;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = __shfl_xor_sync(ARX_FULL_MASK, best, o);
    best = other > best ? other : best;
  }
  if ((threadIdx.x & 31) == 0) atomicMax(packed, best);
}

__device__ __forceinline__ float unpack_value(unsigned long long p) {
  unsigned int u = (unsigned int)(p >> 32);
  u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
  return __uint_as_float(u);
}

// one block per catalog item v, threads over the batch columns b
__global__ void token_pool_fwd_kernel(const float* __restrict__ S, const float* __restrict__ bias, long long mb,
                                      const int* __restrict__ values, const long long* __restrict__ ptr, long long V,
                                      int mode, const unsigned long long* __restrict__ packed_max, float scale,
                                      float* __restrict__ out, int* __restrict__ argmax, float* __restrict__ denom) {
  const long long v = blockIdx.x;
  const long long p0 = ptr ? ptr[v] : v, p1 = ptr ? ptr[v + 1] : v + 1;
  const float m = (mode == TP_LSE) ? unpack_value(*packed_max) : 0.f;
  for (long long b = threadIdx.x; b < mb; b += blockDim.x) {
    float acc = (mode == TP_MAX) ? -CUDART_INF_F : 0.f;
    int best = -1;
    for (long long p = p0; p < p1; ++p) {
      const int t = __ldg(values + p);
      const float s = S[(long long)t * mb + b] + (bias ? __ldg(bias + t) : 0.f);
      if (mode == TP_MAX) { if (s > acc) { acc = s; best = t; } }           // first maximum wins (segment_max ties)
      else if (mode == TP_LSE) acc += __expf(s - m);
      else acc += s;
    }
    float r = acc;
    if (mode == TP_LSE) { if (denom) denom[v * mb + b] = 1.0f + acc; r = m + __logf(1.0f + acc); }
    if (mode == TP_MAX && argmax) argmax[v * mb + b] = best;
    out[v * mb + b] += scale * r;
  }
}

__global__ void token_pool_bwd_kernel(const float* __restrict__ dOut, const float* __restrict__ S,
                                      const float* __restrict__ bias, long long mb, const int* __restrict__ values,
                                      const long long* __restrict__ ptr, long long V, int mode,
                                      const unsigned long long* __restrict__ packed_max, float scale,
                                      const int* __restrict__ argmax, const float* __restrict__ denom,
                                      float* __restrict__ dS, float* __restrict__ dmax_accum) {
  const long long v = blockIdx.x;
  const long long p0 = ptr ? ptr[v] : v, p1 = ptr ? ptr[v + 1] : v + 1;
  const float m = (mode == TP_LSE) ? unpack_value(*packed_max) : 0.f;
  float gm = 0.f;
  for (long long b = threadIdx.x; b < mb; b += blockDim.x) {
    const float g = scale * dOut[v * mb + b];
    if (mode == TP_MAX) {
      const int t = argmax[v * mb + b];
      if (t >= 0) atomicAdd(dS + (long long)t * mb + b, g);
    } else if (mode == TP_LSE) {
      const float inv = 1.0f / denom[v * mb + b];
      for (long long p = p0; p < p1; ++p) {
        const int t = __ldg(values + p);
        const float s = S[(long long)t * mb + b] + (bias ? __ldg(bias + t) : 0.f);
        atomicAdd(dS + (long long)t * mb + b, g * __expf(s - m) * inv);
      }
      gm += g * inv;                                  // d/dm of m + log(1 + sum exp(s - m)) = 1 / (1 + sum)
    } else {
      for (long long p = p0; p < p1; ++p) atomicAdd(dS + (long long)__ldg(values + p) * mb + b, g);
    }
  }
  if (mode == TP_LSE) {
    gm = warp_sum(gm);
    if ((threadIdx.x & 31) == 0 && gm != 0.f) atomicAdd(dmax_accum, gm);
  }
}

// the gradient collected for the global maximum goes to the element that attains it (tf.reduce_max)
__global__ void token_pool_max_grad_kernel(const unsigned long long* __restrict__ packed_max,
                                           const float* __restrict__ dmax_accum, float* __restrict__ dS) {
  const unsigned int idx = 0xffffffffu - (unsigned int)(*packed_max & 0xffffffffull);
  dS[idx] += *dmax_accum;
}

__global__ void rowsum_kernel(const float* __restrict__ X, long long rows, long long cols, float* __restrict__ out, int accumulate) {
  const int lane = threadIdx.x & 31;
  const long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= rows) return;
  float s = 0.f;
  for (long long c = lane; c < cols; c += 32) s += X[r * cols + c];
  s = warp_sum(s);
  if (lane == 0) out[r] = accumulate ? out[r] + s : s;
}

}  // namespace

extern "C" int arx_score_max(const float* S, const float* bias, int64_t Vf, int64_t mb, uint64_t* packed, void* stream) {
  if (!S || !packed || Vf < 1 || mb < 1) return ARX_E_BADARG;
  if (Vf * mb >= (1ll << 32)) return ARX_E_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(packed, 0, sizeof(uint64_t), st) != cudaSuccess) return ARX_E_LAUNCH;
  score_max_kernel<<<arx_num_sms() * 8, 256, 0, st>>>(S, bias, (long long)Vf, (long long)mb, (unsigned long long*)packed);
  ARX_CHECK_LAUNCH();
  return ARX_OK;
}

extern "C" int arx_token_pool_fwd(const float* S, const float* bias, int64_t mb, const int32_t* values, const int64_t* ptr,
                                  int64_t V, int mode, const uint64_t* packed_max, float scale, float* out,
                                  int32_t* argmax, float* denom, void* stream) {
  if (!S || !values || !out || mb < 1 || V < 0 || mode < TP_GATHER || mode > TP_LSE) return ARX_E_BADARG;
  if (mode == TP_LSE && (!packed_max || !denom)) return ARX_E_BADARG;
  if (mode == TP_MAX && !argmax) return ARX_E_BADARG;
  if (V == 0) return ARX_OK;
  if (V > 2147483647ll) return ARX_E_UNSUPPORTED;
  const int threads = mb >= 256 ? 256 : (mb >= 128 ? 128 : (mb >= 64 ? 64 : 32));
  token_pool_fwd_kernel<<<(unsigned)V, threads, 0, (cudaStream_t)stream>>>(
      S, bias, (long long)mb, values, (const long long*)ptr, (long long)V, mode, (const unsigned long long*)packed_max,
      scale, out, argmax, denom);
  ARX_CHECK_LAUNCH();
  return ARX_OK;
}

extern "C" int arx_token_pool_bwd(const float* dOut, const float* S, const float* bias, int64_t mb, const int32_t* values,
                                  const int64_t* ptr, int64_t V, int mode, const uint64_t* packed_max, float scale,
                                  const int32_t* argmax, const float* denom, float* dS, float* dmax_scratch,
                                  void* stream) {
  if (!dOut || !S || !values || !dS || mb < 1 || V < 0 || mode < TP_GATHER || mode > TP_LSE) return ARX_E_BADARG;
  if (mode == TP_LSE && (!packed_max || !denom || !dmax_scratch)) return ARX_E_BADARG;
  if (mode == TP_MAX && !argmax) return ARX_E_BADARG;
  if (V == 0) return ARX_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (mode == TP_LSE && cudaMemsetAsync(dmax_scratch, 0, sizeof(float), st) != cudaSuccess) return ARX_E_LAUNCH;
  const int threads = mb >= 256 ? 256 : (mb >= 128 ? 128 : (mb >= 64 ? 64 : 32));
  token_pool_bwd_kernel<<<(unsigned)V, threads, 0, st>>>(dOut, S, bias, (long long)mb, values, (const long long*)ptr,
                                                         (long long)V, mode, (const unsigned long long*)packed_max, scale,
                                                         argmax, denom, dS, dmax_scratch);
  if (mode == TP_LSE)
    token_pool_max_grad_kernel<<<1, 1, 0, st>>>((const unsigned long long*)packed_max, dmax_scratch, dS);
  ARX_CHECK_LAUNCH();
  return ARX_OK;
}

extern "C" int arx_rowsum(const float* X, int64_t rows, int64_t cols, float* out, int accumulate, void* stream) {
  if (!X || !out || rows < 0 || cols < 0) return ARX_E_BADARG;
  if (rows == 0) return ARX_OK;
  rowsum_kernel<<<(unsigned)((rows * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(X, (long long)rows, (long long)cols,
                                                                                      out, accumulate);
  ARX_CHECK_LAUNCH();
  return ARX_OK;
}

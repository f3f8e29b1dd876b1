// Micro-benchmark: what bandwidth can random 512-byte row traffic reach on this GPU?
// Gives the practical ceiling for the embedding kernels (pool_fwd = random row gather,
// pool_bwd_apply = random row read-modify-write of table + accumulator).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_rows microbench_rows.cu
//   ./microbench_rows
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <numeric>
#include <random>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("cuda error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int D = 128;   // floats per row (512 B)

template <int U>
__global__ void gather_kernel(const float* __restrict__ tab, const int* __restrict__ idx, int n, float* out) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  float4 acc = make_float4(0, 0, 0, 0);
  for (long long r0 = warp0 * U; r0 < n; r0 += nwarps * U) {
    float4 v[U];
#pragma unroll
    for (int q = 0; q < U; ++q) {
      const int r = (r0 + q < n) ? __ldg(idx + r0 + q) : 0;
      v[q] = __ldg(reinterpret_cast<const float4*>(tab + (size_t)r * D) + lane);
    }
#pragma unroll
    for (int q = 0; q < U; ++q) { acc.x += v[q].x; acc.y += v[q].y; acc.z += v[q].z; acc.w += v[q].w; }
  }
  if (acc.x == 123.456f) out[0] = acc.y + acc.z + acc.w;
}

// ids loaded one per lane first (one coalesced load per 32 rows), then rows
template <int U>
__global__ void gather2_kernel(const float* __restrict__ tab, const int* __restrict__ idx, int n, float* out) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  float4 acc = make_float4(0, 0, 0, 0);
  for (long long r0 = warp0 * 32; r0 < n; r0 += nwarps * 32) {
    const int mine = (r0 + lane < n) ? __ldg(idx + r0 + lane) : 0;
    for (int b = 0; b < 32; b += U) {
      float4 v[U];
#pragma unroll
      for (int q = 0; q < U; ++q) {
        const int r = __shfl_sync(0xffffffffu, mine, b + q);
        v[q] = __ldg(reinterpret_cast<const float4*>(tab + (size_t)r * D) + lane);
      }
#pragma unroll
      for (int q = 0; q < U; ++q) { acc.x += v[q].x; acc.y += v[q].y; acc.z += v[q].z; acc.w += v[q].w; }
    }
  }
  if (acc.x == 123.456f) out[0] = acc.y + acc.z + acc.w;
}

template <int U>
__global__ void rmw_kernel(float* __restrict__ tab, float* __restrict__ accu, const int* __restrict__ idx, int n, float lr) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r0 = warp0 * 32; r0 < n; r0 += nwarps * 32) {
    const int mine = (r0 + lane < n) ? __ldg(idx + r0 + lane) : -1;
    for (int b = 0; b < 32; b += U) {
      float4 e[U], a[U]; int rr[U];
#pragma unroll
      for (int q = 0; q < U; ++q) {
        rr[q] = __shfl_sync(0xffffffffu, mine, b + q);
        if (rr[q] >= 0) {
          e[q] = reinterpret_cast<const float4*>(tab + (size_t)rr[q] * D)[lane];
          a[q] = reinterpret_cast<const float4*>(accu + (size_t)rr[q] * D)[lane];
        }
      }
#pragma unroll
      for (int q = 0; q < U; ++q) {
        if (rr[q] < 0) continue;
        const float g = 0.001f;
        a[q].x = fmaf(g, g, a[q].x); a[q].y = fmaf(g, g, a[q].y); a[q].z = fmaf(g, g, a[q].z); a[q].w = fmaf(g, g, a[q].w);
        e[q].x = fmaf(-lr * g, rsqrtf(a[q].x), e[q].x); e[q].y = fmaf(-lr * g, rsqrtf(a[q].y), e[q].y);
        e[q].z = fmaf(-lr * g, rsqrtf(a[q].z), e[q].z); e[q].w = fmaf(-lr * g, rsqrtf(a[q].w), e[q].w);
        reinterpret_cast<float4*>(accu + (size_t)rr[q] * D)[lane] = a[q];
        reinterpret_cast<float4*>(tab + (size_t)rr[q] * D)[lane] = e[q];
      }
    }
  }
}

// streaming read-modify-write of the same byte volume (rows in order): the DRAM ceiling for RMW
__global__ void rmw_seq_kernel(float* __restrict__ tab, float* __restrict__ accu, long long nvec, float lr) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    float4 e = reinterpret_cast<float4*>(tab)[i], a = reinterpret_cast<float4*>(accu)[i];
    const float g = 0.001f;
    a.x = fmaf(g, g, a.x); a.y = fmaf(g, g, a.y); a.z = fmaf(g, g, a.z); a.w = fmaf(g, g, a.w);
    e.x = fmaf(-lr * g, rsqrtf(a.x), e.x); e.y = fmaf(-lr * g, rsqrtf(a.y), e.y);
    e.z = fmaf(-lr * g, rsqrtf(a.z), e.z); e.w = fmaf(-lr * g, rsqrtf(a.w), e.w);
    reinterpret_cast<float4*>(accu)[i] = a; reinterpret_cast<float4*>(tab)[i] = e;
  }
}

__global__ void flush_kernel(float* buf, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) buf[i] += 1.f;
}

template <typename F>
float time_us(F launch, float* flushbuf, long long nflush, int reps = 5) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int r = 0; r < reps + 1; ++r) {
    flush_kernel<<<148 * 8, 256>>>(flushbuf, nflush);
    CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (r > 0) best = std::min(best, ms * 1e3f);
  }
  return best;
}

int main() {
  const long long V = 2000000;            // 2 M rows x 512 B = 1 GB per table
  const int n = 250000;                   // distinct rows touched per launch (like one C2 batch side)
  float *tab, *accu, *out, *fl; int* idx;
  CK(cudaMalloc(&tab, V * D * 4)); CK(cudaMalloc(&accu, V * D * 4)); CK(cudaMalloc(&out, 64));
  const long long nflush = 64ll << 20;    // 256 MB > L2
  CK(cudaMalloc(&fl, nflush * 4)); CK(cudaMemset(fl, 0, nflush * 4));
  CK(cudaMemset(tab, 0, V * D * 4)); CK(cudaMemset(accu, 0x3f, V * D * 4));
  std::vector<int> perm(V); std::iota(perm.begin(), perm.end(), 0);
  std::mt19937 rng(1); std::shuffle(perm.begin(), perm.end(), rng);
  CK(cudaMalloc(&idx, n * 4)); CK(cudaMemcpy(idx, perm.data(), n * 4, cudaMemcpyHostToDevice));
  int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  printf("SMs %d, rows %d x %d B\n", sms, n, D * 4);
  const double gbytes = (double)n * D * 4 / 1e9;

  { // sequential references
    float us = time_us([&] { CK(cudaMemcpyAsync(accu, tab, (size_t)n * D * 4 * 4, cudaMemcpyDeviceToDevice)); }, fl, nflush);
    printf("memcpy D2D %.0f MB              : %7.1f us  %7.1f GB/s (read+write)\n", gbytes * 4e3, us, 2 * gbytes * 4 / us * 1e6);
    us = time_us([&] { rmw_seq_kernel<<<sms * 8, 256>>>(tab, accu, (long long)n * D / 4, 0.1f); }, fl, nflush);
    printf("sequential RMW (E+A, %d rows)   : %7.1f us  %7.1f GB/s\n", n, us, 4 * gbytes / us * 1e6);
  }
  for (int cps : {2, 4, 8}) {
    float us;
    us = time_us([&] { gather_kernel<4><<<sms * cps, 256>>>(tab, idx, n, out); }, fl, nflush);
    printf("gather  U=4  %d CTA/SM           : %7.1f us  %7.1f GB/s\n", cps, us, gbytes / us * 1e6);
    us = time_us([&] { gather_kernel<8><<<sms * cps, 256>>>(tab, idx, n, out); }, fl, nflush);
    printf("gather  U=8  %d CTA/SM           : %7.1f us  %7.1f GB/s\n", cps, us, gbytes / us * 1e6);
    us = time_us([&] { gather2_kernel<8><<<sms * cps, 256>>>(tab, idx, n, out); }, fl, nflush);
    printf("gather2 U=8  %d CTA/SM           : %7.1f us  %7.1f GB/s\n", cps, us, gbytes / us * 1e6);
    us = time_us([&] { gather2_kernel<16><<<sms * cps, 256>>>(tab, idx, n, out); }, fl, nflush);
    printf("gather2 U=16 %d CTA/SM           : %7.1f us  %7.1f GB/s\n", cps, us, gbytes / us * 1e6);
  }
  for (int cps : {2, 4, 8}) {
    float us;
    us = time_us([&] { rmw_kernel<2><<<sms * cps, 256>>>(tab, accu, idx, n, 0.1f); }, fl, nflush);
    printf("RMW U=2 %d CTA/SM                : %7.1f us  %7.1f GB/s\n", cps, us, 4 * gbytes / us * 1e6);
    us = time_us([&] { rmw_kernel<4><<<sms * cps, 256>>>(tab, accu, idx, n, 0.1f); }, fl, nflush);
    printf("RMW U=4 %d CTA/SM                : %7.1f us  %7.1f GB/s\n", cps, us, 4 * gbytes / us * 1e6);
    us = time_us([&] { rmw_kernel<8><<<sms * cps, 256>>>(tab, accu, idx, n, 0.1f); }, fl, nflush);
    printf("RMW U=8 %d CTA/SM                : %7.1f us  %7.1f GB/s\n", cps, us, 4 * gbytes / us * 1e6);
  }
  // size sweep for the gather: launch ramp vs steady state
  for (int nn : {25000, 50000, 100000, 250000}) {
    float us = time_us([&] { gather2_kernel<8><<<sms * 8, 256>>>(tab, idx, nn, out); }, fl, nflush);
    printf("gather2 U=8 8 CTA/SM rows=%6d : %7.1f us  %7.1f GB/s\n", nn, us, (double)nn * D * 4 / 1e9 / us * 1e6);
    us = time_us([&] { rmw_kernel<4><<<sms * 4, 256>>>(tab, accu, idx, nn, 0.1f); }, fl, nflush);
    printf("RMW U=4 4 CTA/SM rows=%6d     : %7.1f us  %7.1f GB/s\n", nn, us, 4.0 * nn * D * 4 / 1e9 / us * 1e6);
  }
  return 0;
}
